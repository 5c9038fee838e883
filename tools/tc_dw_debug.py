import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
M, co, ci = 64, 128, 64
dY = torch.ones(M, co, device=dev); P = torch.ones(M, ci, device=dev)
out = ops.tc_dw(dY, co, P, ci, M, co, ci); torch.cuda.synchronize()
print("ones: max", out.max().item(), "min", out.min().item(), "nonzero", int((out != 0).sum()), "of", out.numel())
print(out[:4, :8])
# one-hot rows / columns
dY = torch.zeros(M, co, device=dev); dY[:, 5] = 1.0
P = torch.zeros(M, ci, device=dev); P[:, 3] = 2.0
out = ops.tc_dw(dY, co, P, ci, M, co, ci); torch.cuda.synchronize()
nz = (out != 0).nonzero()
print("one-hot: nonzero positions", nz[:10].tolist(), "values", out[out != 0][:10].tolist(), "expected (5,3)=", 2.0 * M)
# point dependence: only point 0 nonzero
dY = torch.zeros(M, co, device=dev); dY[0, :] = torch.arange(co, device=dev).float()
P = torch.zeros(M, ci, device=dev); P[0, :] = torch.arange(ci, device=dev).float() + 1
out = ops.tc_dw(dY, co, P, ci, M, co, ci); torch.cuda.synchronize()
ref = dY.t() @ P
print("point0 outer: err", (out - ref).abs().max().item(), "ref max", ref.max().item())
print(out[:3, :6]); print(ref[:3, :6])
