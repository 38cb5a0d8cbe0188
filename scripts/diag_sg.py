"""GPU diagnostic: bitwise agreement of our SG kernels with the reference's (oracle/_ref/sgutilslib.so)."""
import importlib.util, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_sg import make
from goliath_b200 import sgutilslib as ours
spec = importlib.util.spec_from_file_location("sgutilslib", os.path.join(ROOT, "oracle/_ref/sgutilslib.so"))
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
dev = torch.device("cuda:0")
res = {}
for wt in range(4):
    N, D, L = 2, 50000, 32
    dirs, sig, lv, lp, pp, nl = make(N=N, D=D, L=L, seed=9)
    if wt >= 2: sig = sig * 0 + 0.3
    a = [t.to(dev).contiguous() for t in (dirs, sig, lv, lp, pp, nl)]
    o = [torch.empty(N, D, 3, device=dev) for _ in range(2)]
    ref.evaluate_gaussian_fwd(*a, o[0], wt); ours.evaluate_gaussian_fwd(*a, o[1], wt)
    g = torch.randn(N, D, 3, device=dev)
    outs = []
    for lib in (ref, ours):
        gd, gs, gl = torch.zeros(N, D, 3, device=dev), torch.zeros(N, D, device=dev), torch.zeros(N, L, 3, device=dev)
        lib.evaluate_gaussian_bwd(*a, g, gd, gs, gl, wt); outs.append((gd, gs, gl))
    torch.cuda.synchronize()
    def st(x, y):
        x, y = x.cpu().numpy(), y.cpu().numpy()
        return dict(bit_equal=float((x.view(np.uint32) == y.view(np.uint32)).mean()),
                    rel=float(np.linalg.norm(x.astype(np.float64) - y) / max(np.linalg.norm(y.astype(np.float64)), 1e-30)))
    res[wt] = dict(fwd=st(o[1], o[0]), gdir=st(outs[1][0], outs[0][0]), gsig=st(outs[1][1], outs[0][1]), glv=st(outs[1][2], outs[0][2]))
print(json.dumps(res, indent=1))
