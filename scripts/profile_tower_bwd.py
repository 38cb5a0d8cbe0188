"""Profiling driver: training forward + backward of the RGCA vnocond tower (256 -> ... -> 125 @ 1024^2), SIMT path.
Usage (launch list): ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_tower_bwd.py
Without ncu it prints CUDA-event times of forward and forward+backward."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from goliath_b200 import nn as gnn

dev = torch.device("cuda:0")
torch.manual_seed(0)
plan = [256, 256, 128, 128, 64, 32, 16, 125]
layers, size = [], 8
for i, (a, b) in enumerate(zip(plan[:-1], plan[1:])):
    size *= 2
    layers += gnn.make_conv_trans(a, b, 4, 2, 1, "wn", torch.nn.LeakyReLU(0.2) if i < 6 else None, ub=(size, size))
tower = torch.nn.Sequential(*layers).to(dev)
x = torch.randn(1, 256, 8, 8, device=dev, requires_grad=True)
go = torch.ones(1, 125, 1024, 1024, device=dev)


def step():
    y = tower(x)
    y.backward(go)
    for p in tower.parameters():
        p.grad = None
    x.grad = None


for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
print("fwd+bwd ms", sorted(ts)[1])
