"""GPU parity of the fused deconv layer (csrc/deconv_wnub.cu, rows R1/R8) and LinearWN against vectors produced by
IMPORTING the reference's own layers (tests/golden/nn_layers_ref.npz <- tests/golden/make_layers_golden.py)."""
import os

import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nn_layers_ref.npz")


def test_deconv_tower_vs_reference_layers(cuda):
    from goliath_b200 import nn as gnn

    g = np.load(GOLD)
    f = lambda k: torch.from_numpy(g[k]).float().to(cuda)
    l1 = gnn.make_conv_trans(5, 11, 4, 2, 1, "wn", torch.nn.LeakyReLU(0.2), ub=(12, 20))
    l2 = gnn.make_conv_trans(11, 7, 4, 2, 1, "wn", ub=(24, 40))
    net = torch.nn.Sequential(*l1, *l2).to(cuda)
    assert len(net) == 3 and isinstance(net[1], gnn.FusedLeakyReLU), "LeakyReLU is fused into the layer, indices kept"
    # the reference's Sequential state-dict names ([deconv, LeakyReLU, deconv] -> 0.*, 2.*) load directly
    sd = {"0.weight_v": f("dc_p0"), "0.weight_g": f("dc_p1"), "0.bias": f("dc_p2"),
          "2.weight_v": f("dc_p3"), "2.weight_g": f("dc_p4"), "2.bias": f("dc_p5")}
    net.load_state_dict(sd)
    x = f("dc_x").requires_grad_()
    y = net(x)
    assert_close(t2n(y), g["dc_y"], rtol=1e-4, atol=1e-5 * np.abs(g["dc_y"]).max(), what="deconv tower output")
    (y * f("dc_w")).sum().backward()
    assert_close(t2n(x.grad), g["dc_gx"], rtol=2e-4, atol=2e-5 * np.abs(g["dc_gx"]).max(), what="grad input")
    params = [net[0].weight_v, net[0].weight_g, net[0].bias, net[2].weight_v, net[2].weight_g, net[2].bias]
    for i, p in enumerate(params):
        r = g["dc_gp%d" % i]
        assert_close(t2n(p.grad), r, rtol=5e-4, atol=5e-5 * np.abs(r).max(), what="grad param %d" % i)


@pytest.mark.parametrize("shape", [(1, 16, 125, 64, 48), (2, 264, 40, 8, 8), (1, 3, 5, 17, 33)])
def test_deconv_kernel_vs_torch(cuda, shape):
    """Odd sizes, partial channel blocks, partial tiles; oracle = torch fp64 conv_transpose2d on the CPU."""
    from goliath_b200 import nn as gnn

    B, Cin, Cout, Hi, Wi = shape
    gen = torch.Generator().manual_seed(Cin * 7 + Cout)
    layer = gnn.ConvTranspose2dWNUB(Cin, Cout, 2 * Hi, 2 * Wi)
    with torch.no_grad():
        layer.weight_v.copy_(torch.randn(layer.weight_v.shape, generator=gen) * 0.1)
        layer.weight_g.copy_(torch.rand(layer.weight_g.shape, generator=gen) + 0.5)
        layer.bias.copy_(torch.randn(layer.bias.shape, generator=gen))
    x = torch.randn(B, Cin, Hi, Wi, generator=gen)
    w = (layer.weight_g * layer.weight_v / layer.weight_v.norm()).double()
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv_transpose2d(x.double(), w, None, 2, 1) + layer.bias.double()[None], 0.2)
    y = layer.to(cuda)(x.to(cuda), slope=0.2)
    assert_close(t2n(y), ref.detach().numpy(), rtol=1e-4, atol=1e-5 * float(ref.abs().max()), what="deconv %s" % (shape,))


def test_linear_wn_vs_reference(cuda):
    from goliath_b200 import nn as gnn

    g = np.load(GOLD)
    f = lambda k: torch.from_numpy(g[k]).float().to(cuda)
    lin = gnn.LinearWN(9, 13).to(cuda)
    lin.load_state_dict({"weight_v": f("li_v"), "weight_g": f("li_g"), "bias": f("li_b")})
    x = f("li_x").requires_grad_()
    y = lin(x)
    assert_close(t2n(y), g["li_y"], rtol=1e-4, atol=1e-5, what="LinearWN")
    (y * f("li_w")).sum().backward()
    assert_close(t2n(x.grad), g["li_gx"], rtol=2e-4, atol=1e-5, what="LinearWN grad x")
    assert_close(t2n(lin.weight_v.grad), g["li_gv"], rtol=5e-4, atol=1e-5, what="LinearWN grad v")
    assert_close(t2n(lin.weight_g.grad), g["li_gg"], rtol=5e-4, atol=1e-5, what="LinearWN grad g")
