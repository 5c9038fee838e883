"""Per-shape timing of the tensor-core kernels (CUDA events, median of 20, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops, _lib

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 120576
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us

print("M =", M)
for (N, K) in [(64, 32), (64, 64), (128, 64), (128, 128), (256, 128), (64, 128), (128, 256)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.2; b = torch.randn(N, device=dev)
    sc, sh = torch.rand(K, device=dev), torch.randn(K, device=dev)
    gamma, beta = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
    rm, rv, nbt = torch.zeros(N, device=dev), torch.ones(N, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    fold = (gamma, beta, 1e-5, rm, rv, nbt, 0.1)
    y, mean, var, scale, shift = ops.tc_gemm(A, K, W, K, False, M, N, K, bias=b, a_aff=(sc, sh, True), stats=True, fold=fold)
    t = timeit(lambda: ops.tc_gemm(A, K, W, K, False, M, N, K, bias=b, a_aff=(sc, sh, True), stats=True, fold=fold))
    t2 = timeit(lambda: ops.tc_gemm(A, K, W, K, False, M, N, K, bias=b))
    G = torch.randn(M, N, device=dev)
    s12 = ops.act_bwd_reduce(G, N, y, N, scale, shift, mean, var, 1e-5, True, M, N)
    mA, vA = A.mean(0), A.var(0, unbiased=False)
    bw = (y, N, scale, shift, True, mean, var, s12, 1e-5)
    tb = timeit(lambda: ops.tc_gemm(G, N, W, K, True, M, K, N, bnbwd=bw + (True,), bnred=(A, K, sc, sh, mA, vA, 1e-5, True)))
    tb1 = timeit(lambda: ops.tc_gemm(G, N, W, K, True, M, K, N, bnbwd=bw + (False,)))
    tb2 = timeit(lambda: ops.tc_gemm(G, N, W, K, True, M, K, N, bnbwd=bw + (True,)))
    tb3 = timeit(lambda: ops.tc_gemm(G, N, W, K, True, M, K, N, bnred=(A, K, sc, sh, mA, vA, 1e-5, True)))
    tb4 = timeit(lambda: ops.tc_gemm(G, N, W, K, True, M, K, N))
    tb5 = timeit(lambda: ops.tc_gemm(A, K, W, K, False, M, N, K, bias=b, stats=True, fold=fold))
    ts = timeit(lambda: ops.gemm(A, K, True, W, K, True, M, N, K, bias=b, a_aff=(sc, sh, True), stats=False))
    fl = 2.0 * M * N * K
    by = 4.0 * M * (N + K)
    print("N=%3d K=%3d: fwd affine+stats %6.1f us (%5.1f TF/s, %4.2f TB/s) | stats only %6.1f | plain %6.1f || bwd dX[M,%d]: all %6.1f | bnbwd %6.1f | bnbwd+dY %6.1f | bnred %6.1f | plain %6.1f || simt fwd %6.1f"
          % (N, K, t, fl / t / 1e6, by / t / 1e6, tb5, t2, K, tb, tb1, tb2, tb3, tb4, ts))
for (co, ci) in [(128, 64), (128, 128), (256, 128), (256, 64)]:
    dY = torch.randn(M, co, device=dev); P = torch.randn(M, ci, device=dev)
    sc, sh = torch.rand(ci, device=dev), torch.randn(ci, device=dev)
    t = timeit(lambda: ops.tc_dw(dY, co, P, ci, M, co, ci, p_aff=(sc, sh, True)))
    ts = timeit(lambda: ops.gemm(dY, co, False, P, ci, False, co, ci, M, b_aff=(sc, sh, True)))
    fl = 2.0 * M * co * ci
    by = 4.0 * M * (co + ci)
    print("dW co=%3d ci=%3d: tc %7.1f us (%6.1f TF/s, %5.2f TB/s) | simt %7.1f us" % (co, ci, t, fl / t / 1e6, by / t / 1e6, ts))
# elementwise / reductions at [M,128]
C = 128
Y = torch.randn(M, C, device=dev); G = torch.randn(M, C, device=dev)
mean, var = ops.colstats(Y, C, M, C)
scale, shift = ops.bn_fold(mean, var, None, None, 1e-5)
t = timeit(lambda: ops.act_bwd_reduce(G, C, Y, C, scale, shift, mean, var, 1e-5, True, M, C))
print("act_bwd_reduce [M,128]: %.1f us (%.2f TB/s)" % (t, 8.0 * M * C / t / 1e6))
s12 = ops.act_bwd_reduce(G, C, Y, C, scale, shift, mean, var, 1e-5, True, M, C)
s1, s2 = s12[:C], s12[C:]
t = timeit(lambda: ops.act_bwd_apply(G, C, Y, C, scale, shift, mean, var, 1e-5, True, True, s1, s2, M, C))
print("act_bwd_apply  [M,128]: %.1f us (%.2f TB/s)" % (t, 12.0 * M * C / t / 1e6))
