"""Generate tests/golden/raymarch_pytorch_ref.npz from the REFERENCE's own numerical oracle.

The only reference-authored oracle on the MVP path is the pure-PyTorch raymarcher written inline in
`gradcheck()` of /root/reference/extensions/mvpraymarch/mvpraymarch.py (:596-652, dense over all primitives,
F.grid_sample(align_corners=True), accum == 0).  This script does NOT copy that code into the repo: at run time it
parses the reference file, takes the body of `gradcheck` up to the "run cuda version" marker, rewrites the device
strings to CPU and the problem size to something tiny, executes it, and stores inputs + outputs (+ autograd
gradients w.r.t. the effective tensors the CUDA entry point receives).  Needs /root/reference, so it runs in the
build container only; the .npz it writes is committed and is what tests/test_oracle_raymarch.py reads.

Usage: python tests/golden/make_raymarch_golden.py
"""
import ast
import os
import re
import sys
import time
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference/extensions/mvpraymarch/mvpraymarch.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raymarch_pytorch_ref.npz")


def reference_body():
    src = open(REF).read()
    tree = ast.parse(src)
    lines = src.split("\n")
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "gradcheck"][0]
    rod = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Rodrigues"][0]
    body = lines[fn.body[0].lineno - 1: fn.end_lineno]
    stop = [i for i, l in enumerate(body) if "run cuda version" in l][0]
    body = body[:stop]
    # cut at the reference's own backward call; gradients are taken below w.r.t. the effective tensors
    cut = [i for i, l in enumerate(body) if "sample0.backward" in l][0]
    body = body[:cut]
    text = "\n".join(l[4:] if l.startswith("    ") else l for l in body)
    text = text.replace('.to("cuda")', "").replace('device="cuda"', 'device="cpu"').replace("torch.cuda.synchronize()", "pass")
    rod_src = "\n".join(lines[rod.lineno - 1: rod.end_lineno])
    return text, rod_src


def run(dowarp, sizes):
    text, rod_src = reference_body()
    for name, val in sizes.items():
        text, n = re.subn(r"^%s = \d+$" % name, "%s = %d" % (name, val), text, count=1, flags=re.M)
        assert n == 1, name
    ns = dict(torch=torch, F=F, nn=nn, time=time, np=np)
    exec(rod_src, ns)
    ns.update(usebvh="fixedorder", sortprims=False, maxhitboxes=512, synchitboxes=True, dowarp=dowarp, chlast=True,
              fadescale=6.5, fadeexp=7.5, accum=0, termthresh=0.0, algo=1 if dowarp else 0, griddim=3,
              blocksize=(8, 16), bwdblocksize=(8, 16))
    exec(text, ns)
    eff = [ns["template"], ns["primpos"], ns["primrot"], ns["primscale"]] + ([ns["warp"]] if dowarp else [])
    grads = torch.autograd.grad(ns["sample0"].sum(), eff, allow_unused=True)
    tag = "w" if dowarp else "n"
    out = {
        tag + "_raypos": ns["_raypos"], tag + "_raydir": ns["_raydir"], tag + "_tminmax": ns["_tminmax"],
        tag + "_stepsize": torch.tensor(ns["_stepsize"]),
        tag + "_template": ns["template"].permute(0, 1, 3, 4, 5, 2), tag + "_primpos": ns["primpos"],
        tag + "_primrot": ns["primrot"], tag + "_primscale": ns["primscale"],
        tag + "_rayrgba": ns["sample0"],
        tag + "_g_template": grads[0].permute(0, 1, 3, 4, 5, 2), tag + "_g_primpos": grads[1],
        tag + "_g_primrot": grads[2], tag + "_g_primscale": grads[3],
    }
    if dowarp:
        out[tag + "_warp"] = ns["warp"].permute(0, 1, 3, 4, 5, 2)
        out[tag + "_g_warp"] = grads[4].permute(0, 1, 3, 4, 5, 2)
    return {k: v.detach().contiguous().numpy().astype(np.float32) for k, v in out.items()}


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("needs /root/reference (build container only)")
    sizes = dict(N=1, H=16, W=20, k3=2, M=8)
    d = {}
    d.update(run(False, sizes))
    d.update(run(True, sizes))
    d["fadescale"] = np.float32(6.5)
    d["fadeexp"] = np.float32(7.5)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;",
          "alpha range", float(d["n_rayrgba"][..., 3].min()), float(d["n_rayrgba"][..., 3].max()))
