"""A/B of the vector-filter ECC kernels (warp-per-node vs stream), L2 flushed before every launch and warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops
from superpoint_graph_b200.synthetic import make_batch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
H = 32


def timed(fn, cold, reps=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


for n, k in ((100000, 8), (300000, 8), (300000, 18)):
    b = make_batch(n_nodes=n, k=k, seed=5, npts=1, minpts=1)
    N, E = b["degs"].numel(), b["idxn"].numel()
    graph = ops.EccGraph(b["idxn"], None, b["degs"], n_in=N)
    x, g, w = torch.randn(N, H, device=dev), torch.randn(N, H, device=dev), torch.randn(E, H, device=dev)
    gw = torch.empty_like(w)
    by = 4 * H * E + 8 * H * N + 4 * E + 4 * (N + 1)
    # plain streaming copy of the same byte count as a yardstick
    src = torch.empty(by // 8, dtype=torch.float32, device=dev); dst = torch.empty_like(src)
    tc = timed(lambda: dst.copy_(src), True)
    row = ["N=%d E=%d  copy(%d MB r+w) %.1f us" % (N, E, by >> 20, tc)]
    for name, thr in (("node", 1 << 60), ("stream", 0)):
        ops.STREAM_MIN_EDGES[0] = thr
        for cold in (True, False):
            tf = timed(lambda: ops.ecc_fwd(x, w, graph, H), cold)
            tb = timed(lambda: ops.ecc_bwd_x(w, g, graph, H), cold)
            row.append("%s/%s fwd %.1f us (%.2f) bwd_x %.1f us (%.2f)" % (
                name, "cold" if cold else "warm", tf, by / tf / 1e3 / 6572.2, tb, (by + 4 * E + 4 * N) / tb / 1e3 / 6572.2))
    tw = timed(lambda: ops.ecc_bwd_w(x, g, graph, (E, H), out=gw), True)
    row.append("bwd_w cold %.1f us (%.2f)" % (tw, by / tw / 1e3 / 6572.2))
    print(" | ".join(row), flush=True)
