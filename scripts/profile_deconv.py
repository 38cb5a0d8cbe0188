"""GPU profiling driver for the decoder's last (bandwidth-bound) layer: ConvTranspose2dWNUB 16 -> 125 @ 512^2 -> 1024^2
through the C ABI (what ncu attaches to); prints CUDA-event times.  Usage: python scripts/profile_deconv.py [Cin Cout Hi reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from goliath_b200 import _lib

Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 125
Hi = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, Cin, Hi, Hi, device=dev)
v = torch.randn(Cin, Cout, 4, 4, device=dev) * 0.1
scale = torch.rand(Cout, device=dev) + 0.5
bias = torch.randn(Cout, 2 * Hi, 2 * Hi, device=dev)
out = torch.empty(1, Cout, 2 * Hi, 2 * Hi, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib(); st = _lib.stream_ptr(dev)
def run():
    _lib.check(L.gb_deconv4x4s2_wnub_fwd(1, Cin, Cout, Hi, Hi, x.data_ptr(), v.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                         0.2, 1, out.data_ptr(), st), "fwd")
ts = []
for _ in range(reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
byts = (2 * Cout * 4 * Hi * Hi + Cin * Hi * Hi) * 4
print("deconv %d->%d @%d^2: %s ms; best %.3f ms = %.0f GB/s algorithmic, %.1f TFLOP/s" % (
    Cin, Cout, Hi, ["%.3f" % t for t in ts], min(ts), byts / min(ts) / 1e6, 2 * 4 * Cin * Cout * 4 * Hi * Hi / min(ts) / 1e9))
ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv_transpose2d(x[:, :, :64, :64].double(), (v * scale.view(1, -1, 1, 1)).double(), None, 2, 1)[:, :, :100, :100] + bias[:, :100, :100].double(), 0.2)
print("max abs err vs torch fp64 (crop):", float((out[:, :, :100, :100].double() - ref).abs().max()))
