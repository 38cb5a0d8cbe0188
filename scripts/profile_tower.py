"""Profiling driver: one inference forward of the RGCA vnocond tower (256 -> ... -> 125 @ 1024^2) on the tensor-core path.
Usage (launch list): ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_tower.py"""
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
x = torch.randn(1, 256, 8, 8, device=dev)
if len(sys.argv) > 1:
    gnn.TC_MIN_CIN = int(sys.argv[1])
with torch.no_grad():
    for _ in range(2):
        y = gnn.tower_forward_tc(tower, x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    y = gnn.tower_forward_tc(tower, x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok", tuple(y.shape))
