import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import _lib
dev = torch.device("cuda:0")
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
sc, sh = torch.rand(K, device=dev), torch.randn(K, device=dev)
img = torch.empty(2 * N * K, device=dev)
_lib.call("spg_tc_pack_weights", W, K, 0, N, K, K, img, _lib.current_stream())
out = torch.empty(M, N, device=dev)
sws = torch.empty((4 * ((M + 127) // 128) + 64) * N * 3, device=dev)
for _ in range(3):
    _lib.call("spg_tc_gemm", A, K, img, b, out, N, M, N, K, sc, sh, 1, sws, _lib.current_stream())
torch.cuda.synchronize()
# weight gradient kernel too
from superpoint_graph_b200 import ops
dY = torch.randn(M, 256, device=dev); P = torch.randn(M, 128, device=dev)
for _ in range(2):
    ops.tc_dw(dY, 256, P, 128, M, 256, 128, p_aff=(sc[:128] if K >= 128 else None, None, True))
torch.cuda.synchronize()
print("done")
