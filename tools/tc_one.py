"""One launch of each tensor-core kernel at a given shape, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:tc_ -o gpurun_out/tc python tools/tc_one.py 120576 256 128
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops
dev = torch.device("cuda:0")
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.2; b = torch.randn(N, device=dev)
sc, sh = torch.rand(K, device=dev), torch.randn(K, device=dev)
gamma, beta = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
rm, rv, nbt = torch.zeros(N, device=dev), torch.ones(N, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
for _ in range(3):  # forward: BN-apply prologue + statistics + fold
    y, mean, var, scale, shift = ops.tc_gemm(A, K, W, K, False, M, N, K, bias=b, a_aff=(sc, sh, True), stats=True,
                                              fold=(gamma, beta, 1e-5, rm, rv, nbt, 0.1))
torch.cuda.synchronize()
G = torch.randn(M, N, device=dev)
s12 = ops.act_bwd_reduce(G, N, y, N, scale, shift, mean, var, 1e-5, True, M, N)
mA, vA = A.mean(0), A.var(0, unbiased=False)
for _ in range(3):  # backward: BN-backward prologue + dY store + sums of the layer below
    ops.tc_gemm(G, N, W, K, True, M, K, N, bnbwd=(y, N, scale, shift, True, mean, var, s12, 1e-5, True),
                bnred=(A, K, sc, sh, mA, vA, 1e-5, True))
torch.cuda.synchronize()
dY = torch.randn(M, 256, device=dev); P = torch.randn(M, 128, device=dev)
for _ in range(2):
    ops.tc_dw(dY, 256, P, 128, M, 256, 128, p_aff=(sc[:128] if K >= 128 else None, None, True))
torch.cuda.synchronize()
print("done")
