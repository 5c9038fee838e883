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
for (N, K) in [(64, 64), (128, 64), (128, 128), (256, 128), (64, 128), (128, 256)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    sc, sh = torch.rand(K, device=dev), torch.randn(K, device=dev)
    img = torch.empty(2 * N * K, device=dev)
    _lib.call("spg_tc_pack_weights", W, K, 0, N, K, K, img, _lib.current_stream())
    out = torch.empty(M, N, device=dev)
    tiles = 4 * ((M + 127) // 128)
    sws = torch.empty((tiles + 64) * N * 3, device=dev)
    def run(stats=True, pro=True):
        _lib.call("spg_tc_gemm", A, K, img, b, out, N, M, N, K, sc if pro else None, sh if pro else None, int(pro),
                  sws if stats else None, _lib.current_stream())
    t = timeit(run)
    t2 = timeit(lambda: run(False, False))
    tp = timeit(lambda: _lib.call("spg_tc_pack_weights", W, K, 0, N, K, K, img, _lib.current_stream()))
    ts = timeit(lambda: ops.gemm(A, K, True, W, K, True, M, N, K, bias=b, a_aff=(sc, sh, True), stats=False))
    fl = 2.0 * M * N * K
    by = 4.0 * M * (N + K)
    print("fwd N=%3d K=%3d: tc %7.1f us (%6.1f TF/s, %5.2f TB/s) | no stats/prologue %7.1f us | pack %5.1f us | simt %7.1f us"
          % (N, K, t, fl / t / 1e6, by / t / 1e6, t2, tp, ts))
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
s1, s2 = ops.act_bwd_reduce(G, C, Y, C, scale, shift, mean, var, 1e-5, True, M, C)
t = timeit(lambda: ops.act_bwd_apply(G, C, Y, C, scale, shift, mean, var, 1e-5, True, True, s1, s2, M, C))
print("act_bwd_apply  [M,128]: %.1f us (%.2f TB/s)" % (t, 12.0 * M * C / t / 1e6))
sws = torch.randn((4 * ((M + 127) // 128) + 64) * C * 3, device=dev).abs()
mo, vo = torch.empty(C, device=dev), torch.empty(C, device=dev)
t = timeit(lambda: _lib.call("spg_colstats_merge", sws, 4 * ((M + 127) // 128), C, mo, vo, _lib.current_stream()))
print("colstats_merge %d partials x %d cols: %.1f us" % ((M + 127) // 128, C, t))
