"""Generate tests/golden/rgca_heads_ref.npz from the REFERENCE's own code for row R2: the Gaussian
heads + SH-diffuse + reflection-direction block of `PrimDecoder.forward`
(/root/reference/ca_code/models/rgca.py:506-546) and the final colour compose (:572-575).

Nothing is copied into the repo: at run time the script parses the reference file, takes the statements of
`PrimDecoder.forward` from the "# diffuse sh" comment up to (not including) the `if preconv_envmap is not None:`
branch, executes them on small random inputs with a stub `self`, and stores inputs, outputs and autograd gradients of a
random linear loss.  Needs /root/reference (build container only); the .npz is committed.

Usage: python tests/golden/make_heads_golden.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch as th
import torch.nn.functional as F

REF = "/root/reference/ca_code/models/rgca.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rgca_heads_ref.npz")
OUT_KEYS = ["primpos", "primqvec", "primscale", "opacity", "sigma", "spec_vis", "spec_dnml", "spec_nml", "diff_color",
            "ref_dirs"]


def reference_segment():
    src = open(REF).read()
    lines = src.split("\n")
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PrimDecoder"][0]
    fwd = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward"][0]
    start = [i for i in range(fwd.lineno, fwd.end_lineno) if "# diffuse sh" in lines[i]][0]
    stop = [i for i in range(start, fwd.end_lineno) if lines[i].strip().startswith("if preconv_envmap is not None")][0]
    seg = lines[start:stop]
    indent = len(seg[0]) - len(seg[0].lstrip())
    return "\n".join(l[indent:] for l in seg), (start + 1, stop)


def main():
    if not os.path.exists(REF):
        sys.exit("needs /root/reference (build container only)")
    text, span = reference_segment()
    g = th.Generator().manual_seed(506)
    B, Hs, Ws = 2, 6, 8
    G = Hs * Ws
    n_color, n_mono = 16, 65
    n_diff = 3 * n_color + n_mono
    f_vnocond = th.randn(B, n_diff + 12, Hs, Ws, generator=g, dtype=th.float64).requires_grad_()
    f_vnocond.data[:, n_diff + 7:n_diff + 10] *= 3.0          # push some scales outside the [0.1, 20] clamp
    f_vnocond.data[:, n_diff + 11] = f_vnocond.data[:, n_diff + 11] * 2.0 - 2.0  # some sigmas below the 0.01 floor
    f_vcond_planes = th.randn(B, 4, Hs, Ws, generator=g, dtype=th.float64).requires_grad_()
    postex = (100.0 * th.randn(B, 3, Hs, Ws, generator=g, dtype=th.float64)).requires_grad_()
    tn = F.normalize(th.randn(B, 3, Hs, Ws, generator=g, dtype=th.float64), dim=1).requires_grad_()
    albedo = th.rand(1, G, 3, generator=g, dtype=th.float64).requires_grad_()
    light_sh = th.randn(B, 3, 81, generator=g, dtype=th.float64)
    campos = 1000.0 * F.normalize(th.randn(B, 3, generator=g, dtype=th.float64), dim=1)

    self = types.SimpleNamespace(n_diff_coeffs=n_diff, n_color_sh_coeffs=n_color, n_mono_sh_coeffs=n_mono, albedo=albedo)
    ns = dict(th=th, F=F, self=self, B=B, f_vnocond=f_vnocond,
              f_vcond=f_vcond_planes.permute(0, 2, 3, 1).reshape(B, -1, 4),
              primposbase=postex.permute(0, 2, 3, 1).reshape(B, -1, 3),
              primnmlbase=tn.permute(0, 2, 3, 1).reshape(B, -1, 3),
              headrel_light_sh=light_sh, headrel_campos=campos)
    exec(text, ns)
    outs = {k: ns[k] for k in OUT_KEYS}
    outs["primscale_preclip"] = outs.pop("primscale")
    primscale_range = [0.1, 20.0]   # rgca.py:47
    outs["primscale"] = outs["primscale_preclip"].clamp(*primscale_range)
    ws = {k: th.randn(v.shape, generator=g, dtype=th.float64) for k, v in outs.items()}
    loss = sum((ws[k] * v).sum() for k, v in outs.items())
    leaves = [f_vnocond, f_vcond_planes, postex, tn, albedo]
    grads = th.autograd.grad(loss, leaves)
    d = {"in_f_vnocond": f_vnocond, "in_f_vcond": f_vcond_planes, "in_postex": postex, "in_tn": tn, "in_albedo": albedo,
         "in_light_sh": light_sh, "in_campos": campos}
    d.update({"out_" + k: v for k, v in outs.items()})
    d.update({"w_" + k: v for k, v in ws.items()})
    d.update({"g_" + n: v for n, v in zip(["f_vnocond", "f_vcond", "postex", "tn", "albedo"], grads)})
    np.savez_compressed(OUT, **{k: v.detach().numpy() for k, v in d.items()})
    print("executed reference lines %d-%d; wrote %s (%d bytes)" % (span[0], span[1], OUT, os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
