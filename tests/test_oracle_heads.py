"""CPU: pin the torch restatement of row R2 (oracle/heads_oracle.py) against vectors produced by executing the
reference's own PrimDecoder.forward source lines (tests/golden/make_heads_golden.py)."""
import os

import numpy as np
import torch

from util import assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rgca_heads_ref.npz")


def test_heads_oracle_matches_reference_lines():
    from oracle import heads_oracle as ho

    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k]).requires_grad_(k in ("in_f_vnocond", "in_f_vcond", "in_postex", "in_tn", "in_albedo"))
    ins = {k[3:]: t(k) for k in g.files if k.startswith("in_")}
    out = ho.gaussian_heads(ins["f_vnocond"], ins["f_vcond"], ins["postex"], ins["tn"], ins["albedo"], ins["light_sh"],
                            ins["campos"])
    keys = [k[4:] for k in g.files if k.startswith("out_")]
    assert len(keys) == 11
    for k in keys:
        assert_close(out[k].detach().numpy(), g["out_" + k], rtol=1e-12, atol=1e-12, what=k)
    loss = sum((torch.from_numpy(g["w_" + k]) * out[k]).sum() for k in keys)
    grads = torch.autograd.grad(loss, [ins[n] for n in ("f_vnocond", "f_vcond", "postex", "tn", "albedo")])
    for n, gr in zip(("f_vnocond", "f_vcond", "postex", "tn", "albedo"), grads):
        assert_close(gr.numpy(), g["g_" + n], rtol=1e-10, atol=1e-12, what="grad " + n)
    # the fixture exercises the clamps
    assert (g["out_primscale"] != g["out_primscale_preclip"]).any() and (g["out_sigma"] == 0.01).any()
