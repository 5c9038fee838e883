"""A few launches of the fused eval-mode PointNet trunk (both chains) for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:pointnet_fused -o gpurun_out/fused python tools/fused_one.py 18248
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import spg_pointnet
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 18248
F = int(sys.argv[2]) if len(sys.argv) > 2 else 14
torch.manual_seed(0)
net = spg_pointnet.PointNet([64, 64, 128, 128, 256], [256, 64, 32], [64, 64, 128], [128, 64], F, F, prelast_do=0).to(dev).eval()
x, xg = torch.randn(B, F, 128, device=dev) * 0.4, torch.rand(B, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = net(x, xg)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
with torch.no_grad():
    ev[0].record()
    for _ in range(5):
        net(x, xg)
    ev[1].record()
torch.cuda.synchronize()
print("PointNet eval forward, %d clouds: %.3f ms" % (B, ev[0].elapsed_time(ev[1]) / 5))
