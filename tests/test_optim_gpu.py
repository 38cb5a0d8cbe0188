"""goliath_b200.optim.FusedAdam against the reference's own train-step sequence (ca_code/utils/train.py:209-215, restated
line for line below) over torch.optim.Adam / AdamW — the optimizers the reference's configs name."""
import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu


def _reference_update(params, opt):
    # ca_code/utils/train.py:209-215
    for p in params:
        if hasattr(p, "grad") and p.grad is not None:
            p.grad.data[torch.isnan(p.grad.data)] = 0
            p.grad.data[torch.isinf(p.grad.data)] = 0
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()


@pytest.mark.parametrize("adamw,wd", [(False, 0.0), (False, 0.01), (True, 0.05)])
def test_fused_adam_matches_reference_sequence(cuda, adamw, wd):
    from goliath_b200.optim import FusedAdam

    gen = torch.Generator().manual_seed(209)
    shapes = [(3,), (17, 5), (125, 64, 64), (40000,), (16385,), (1,)]   # crosses chunk boundaries, odd sizes
    base = [torch.randn(s, generator=gen) for s in shapes]
    pa = [b.clone().to(cuda).requires_grad_() for b in base]
    pb = [b.clone().to(cuda).requires_grad_() for b in base]
    groups = lambda ps: [dict(params=ps[:3], lr=5e-4), dict(params=ps[3:], lr=1e-3)]   # per-module learning rates
    ref = (torch.optim.AdamW if adamw else torch.optim.Adam)(groups(pa), weight_decay=wd)
    ours = FusedAdam(groups(pb), weight_decay=wd, adamw=adamw, max_grad_norm=1.0, write_clipped_grads=True)
    for it in range(4):
        grads = [torch.randn(s, generator=gen) * (10.0 if it % 2 == 0 else 0.01) for s in shapes]   # clipped / not clipped
        grads[2][3, 5, 7] = float("nan")
        grads[2][9, 1, 1] = float("inf")
        grads[4][16384] = float("-inf")
        for a, b, g in zip(pa, pb, grads):
            a.grad = g.clone().to(cuda)
            b.grad = g.clone().to(cuda)
        if it == 2:       # a parameter without a gradient is skipped by both
            pa[1].grad = None
            pb[1].grad = None
        _reference_update(pa, ref)
        ours.step()
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(pa, pb)):
            assert_close(t2n(b), t2n(a), rtol=2e-6, atol=1e-7, what="param %d after step %d" % (i, it))
            if a.grad is not None:
                assert_close(t2n(b.grad), t2n(a.grad), rtol=2e-6, atol=1e-9, what="clipped grad %d" % i)
                assert torch.isfinite(b.grad).all()
    for a, b in zip(pa, pb):
        sa, sb = ref.state[a], ours.state[b]
        assert_close(t2n(sb["exp_avg"]), t2n(sa["exp_avg"]), rtol=1e-5, atol=1e-9, what="exp_avg")
        assert_close(t2n(sb["exp_avg_sq"]), t2n(sa["exp_avg_sq"]), rtol=5e-5, atol=1e-12, what="exp_avg_sq")
    # torch.optim.Adam's state layout: the reference's optimizer checkpoints load
    sd = ref.state_dict()
    fresh = FusedAdam(groups([b.clone().to(cuda).requires_grad_() for b in base]), weight_decay=wd, adamw=adamw)
    fresh.load_state_dict(sd)
    assert fresh._steps == 4


def test_fused_adam_refuses_cpu_parameters():
    from goliath_b200.optim import FusedAdam

    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()
