"""Generate tests/golden/hand_mvp_ref.npz by RUNNING the reference's own code on CPU (fp64):

  * la.Conv2dWNUB chain (ca_code/nn/layers.py:276-327,472) and blocks.ConvBlock (ca_code/nn/blocks.py:232-280), imported;
  * PoseEncoder, TransDecoder head scaling, GeomDecoder's transform composition, axisangle_to_matrix and the
    slab -> primitive re-layout, executed from /root/reference/ca_code/models/hand_mvp.py by slicing the statements
    out of the file's AST at run time (the module itself cannot be imported: drtk / pytorch3d are absent) —
    nothing is copied into this repository;
  * the state_dict key / shape lists of the reference's PoseEncoder / TransDecoder / DeconvContentDecoder.

Needs /root/reference; the .npz is committed.   Usage: python tests/golden/make_hand_mvp_golden.py
"""
import ast
import json
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np
import torch as th

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hand_mvp_ref.npz")
SRC = "/root/reference/ca_code/models/hand_mvp.py"


def _slice(tree, src, cls=None, func=None, first=None, last=None):
    """source text of a top-level class/function, or of the statements [first, last] (line numbers) inside a method"""
    for node in tree.body:
        if cls and isinstance(node, ast.ClassDef) and node.name == cls:
            if func is None:
                return ast.get_source_segment(src, node)
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == func:
                    stmts = [s for s in sub.body if s.lineno >= first and s.end_lineno <= last]
                    return "\n".join(_dedent(ast.get_source_segment(src, s, padded=True)) for s in stmts)
        if not cls and isinstance(node, ast.FunctionDef) and node.name == func:
            return ast.get_source_segment(src, node)
    raise KeyError((cls, func))


def _dedent(text):
    import textwrap
    return textwrap.dedent(text)


def _jitter(mod):
    for m in mod.modules():
        if hasattr(m, "weight_g"):
            m.weight_g.data.mul_(1.0 + 0.3 * th.randn_like(m.weight_g))
        b = getattr(m, "bias", None)
        if b is not None:
            b.data.normal_(0, 0.3)


def _grads(y, w, inputs):
    return th.autograd.grad((y * w).sum(), inputs, allow_unused=True)


def main():
    if not os.path.isfile(SRC):
        sys.exit("needs /root/reference (build container only)")
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    sys.modules.setdefault("turtle", MagicMock())  # blocks.py:8 imports it by accident; tkinter is absent here
    import ca_code.nn.layers as la
    from ca_code.nn.blocks import ConvBlock, tile2d

    src = open(SRC).read()
    tree = ast.parse(src)
    ns = {"th": th, "nn": th.nn, "la": la, "ConvBlock": ConvBlock, "tile2d": tile2d, "Tuple": tuple, "F": th.nn.functional}
    for name in ("PoseEncoder", "TransDecoder", "DeconvContentDecoder"):
        exec(_slice(tree, src, cls=name), ns)
    exec(_slice(tree, src, func="axisangle_to_matrix"), ns)

    th.manual_seed(4471)
    d = {}

    # (1) la.Conv2dWNUB chain 5 -> 11 (LeakyReLU) -> 7 on 12x20
    net = th.nn.Sequential(la.Conv2dWNUB(5, 11, 12, 20, 3, 1, 1), th.nn.LeakyReLU(0.2), la.Conv2dWNUB(11, 7, 12, 20, 3, 1, 1)).double()
    for m in net:
        la.glorot(m, 0.2)
    _jitter(net)
    x = th.randn(2, 5, 12, 20, dtype=th.float64, requires_grad=True)
    y = net(x)
    w = th.randn_like(y)
    params = [net[0].weight_v, net[0].weight_g, net[0].bias, net[2].weight_v, net[2].weight_g, net[2].bias]
    g = _grads(y, w, [x] + params)
    d.update(cv_x=x, cv_y=y, cv_w=w, cv_gx=g[0])
    for i, (p, gp) in enumerate(zip(params, g[1:])):
        d["cv_p%d" % i], d["cv_gp%d" % i] = p, gp
    wv, wg = net[0].weight_v, net[0].weight_g
    assert wg.shape == (11, 1, 1, 1) and th.allclose(net[0].weight, wg * wv / wv.norm(), atol=1e-12)

    # (2) ConvBlock 3x3 (6 -> 10 @9x9) and 1x1 (7 -> 5 @9x9)
    for tag, blk, cin in (("b3", ConvBlock(6, 10, 9), 6), ("b1", ConvBlock(7, 5, 9, kernel_size=1, padding=0), 7)):
        blk = blk.double()
        _jitter(blk)
        xb = th.randn(2, cin, 9, 9, dtype=th.float64, requires_grad=True)
        yb = blk(xb)
        wb = th.randn_like(yb)
        names = [n for n, _ in blk.named_parameters()]
        gb = _grads(yb, wb, [xb] + [p for _, p in blk.named_parameters()])
        d.update({tag + "_x": xb, tag + "_y": yb, tag + "_w": wb, tag + "_gx": gb[0]})
        for n, p, gp in zip(names, [p for _, p in blk.named_parameters()], gb[1:]):
            d["%s_p_%s" % (tag, n)], d["%s_g_%s" % (tag, n)] = p, gp
        d[tag + "_names"] = np.array(json.dumps(names))

    # (3) PoseEncoder (sliced class) small: pose [2, 6+8] -> [2, 12, 6, 6]
    pe = ns["PoseEncoder"](8, 12, 6).double()
    _jitter(pe)
    pose = th.randn(2, 14, dtype=th.float64, requires_grad=True)
    yp = pe(pose)
    wp = th.randn_like(yp)
    d.update(pe_pose=pose, pe_y=yp, pe_w=wp, pe_gpose=_grads(yp, wp, [pose])[0])
    for n, p in pe.named_parameters():
        d["pe_p_" + n] = p

    # (4) TransDecoder head scaling (hand_mvp.py:317-320) + GeomDecoder composition (:417-425)
    B, K = 2, 16
    dec = (3.0 * th.randn(B, 9, 4, 4, dtype=th.float64)).requires_grad_(True)
    q, _ = th.linalg.qr(th.randn(B, K, 3, 3, dtype=th.float64))
    primrotbase = q.contiguous()
    primposbase = th.randn(B, K, 3, dtype=th.float64)
    env = dict(ns)
    env.update(out=dec, local_encoding=dec, self=types.SimpleNamespace(prim_scale=512, training=False, primposstart=1000))
    exec(_slice(tree, src, cls="TransDecoder", func="forward", first=317, last=320), env)
    env.update(delta_pos=env["primposdelta"], delta_rvec=env["primrvecdelta"], delta_scale=env["primscaledelta"], B=B,
               primposbase=primposbase, primrotbase=primrotbase)
    exec(_slice(tree, src, cls="GeomDecoder", func="forward", first=417, last=425), env)
    primpos, primrot, primscale = env["primpos"], env["primrot"], env["primscale"]
    w1, w2, w3 = th.randn_like(primpos), th.randn_like(primrot), th.randn_like(primscale)
    gdec = th.autograd.grad((primpos * w1).sum() + (primrot * w2).sum() + (primscale * w3).sum(), dec)[0]
    d.update(pt_dec=dec, pt_posbase=primposbase, pt_rotbase=primrotbase, pt_primpos=primpos, pt_primrot=primrot,
             pt_primscale=primscale, pt_w1=w1, pt_w2=w2, pt_w3=w3, pt_gdec=gdec)

    # (5) slab -> primitives (hand_mvp.py:172-185): primsize (4,4,2), U=16 -> 4x4 primitives
    me = types.SimpleNamespace(primsize=(4, 4, 2), n_prim_x=4, n_prim_y=4, n_prims=16)
    preds = {"primrgb": th.randn(2, 2, 3, 16, 16, dtype=th.float64), "primalpha": th.randn(2, 2, 1, 16, 16, dtype=th.float64)}
    env = dict(ns)
    env.update(self=me, preds=preds, B=2)
    exec(_slice(tree, src, cls="AutoEncoder", func="render", first=172, last=185), env)
    d.update(sl_rgb=preds["primrgb"], sl_alpha=preds["primalpha"], sl_primrgba=preds["primrgba"])

    # (6) state_dict layouts of the real-size modules
    layout = {}
    for name, mod in (("PoseEncoder", ns["PoseEncoder"](48, 64, 64)), ("TransDecoder", ns["TransDecoder"](64)),
                      ("DeconvContentDecoder", ns["DeconvContentDecoder"](8, 66, 3))):
        layout[name] = [[k, list(v.shape)] for k, v in mod.state_dict().items()]
    d["layouts"] = np.array(json.dumps(layout))

    out = {}
    for k, v in d.items():
        out[k] = v.detach().numpy() if isinstance(v, th.Tensor) else v
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
