"""Generate tests/golden/nn_layers_ref.npz by IMPORTING the reference's own layers
(/root/reference/ca_code/nn/layers.py: ConvTranspose2dWNUB :478-480, LinearWN :468, make_conv_trans :46) and running
them on CPU: a deconv tower shaped like the RGCA ones (k4 s2 p1, weight-norm with whole-tensor norm, untied bias,
LeakyReLU(0.2)) and a LinearWN, forward + autograd gradients.  Needs /root/reference; the .npz is committed.

Usage: python tests/golden/make_layers_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch as th

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nn_layers_ref.npz")


def main():
    if not os.path.isdir("/root/reference/ca_code"):
        sys.exit("needs /root/reference (build container only)")
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    import ca_code.nn.layers as la

    th.manual_seed(331)
    d = {}
    # two chained deconvs: 5 -> 11 (with LeakyReLU) -> 7 (no activation), 6x10 -> 12x20 -> 24x40
    l1 = la.make_conv_trans(5, 11, 4, 2, 1, "wn", th.nn.LeakyReLU(0.2), ub=(12, 20))
    l2 = la.make_conv_trans(11, 7, 4, 2, 1, "wn", ub=(24, 40))
    net = th.nn.Sequential(*l1, *l2).double()
    for m in net:
        la.glorot(m, 0.2)
        if hasattr(m, "bias") and m.bias is not None:
            m.bias.data.normal_(0, 0.3)
        if hasattr(m, "weight_g"):
            m.weight_g.data.mul_(1.0 + 0.3 * th.randn_like(m.weight_g))
    x = th.randn(2, 5, 6, 10, dtype=th.float64, requires_grad=True)
    y = net(x)
    w = th.randn_like(y)
    params = [net[0].weight_v, net[0].weight_g, net[0].bias, net[2].weight_v, net[2].weight_g, net[2].bias]
    grads = th.autograd.grad((y * w).sum(), [x] + params)
    d.update(dc_x=x, dc_y=y, dc_w=w, dc_gx=grads[0])
    for i, (p, g) in enumerate(zip(params, grads[1:])):
        d["dc_p%d" % i] = p
        d["dc_gp%d" % i] = g
    # check the non-standard weight norm the survey describes: w = g * v / ||v||_F
    wv, wg = net[0].weight_v, net[0].weight_g
    assert th.allclose(net[0].weight, wg * wv / wv.norm(), atol=1e-12)
    # LinearWN 9 -> 13
    lin = la.LinearWN(9, 13).double()
    lin.weight_g.data.mul_(1.0 + 0.3 * th.randn_like(lin.weight_g))
    xl = th.randn(4, 9, dtype=th.float64, requires_grad=True)
    yl = lin(xl)
    wl = th.randn_like(yl)
    gl = th.autograd.grad((yl * wl).sum(), [xl, lin.weight_v, lin.weight_g, lin.bias])
    d.update(li_x=xl, li_y=yl, li_w=wl, li_v=lin.weight_v, li_g=lin.weight_g, li_b=lin.bias, li_gx=gl[0], li_gv=gl[1],
             li_gg=gl[2], li_gb=gl[3])
    np.savez_compressed(OUT, **{k: v.detach().numpy() for k, v in d.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes; state_dict keys:", list(net[0].state_dict().keys()))


if __name__ == "__main__":
    main()
