"""ECC kernels at configs[4] scale, for `ncu --set full -k regex:ecc_` captures (profiles/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from superpoint_graph_b200 import ops  # noqa: E402
from superpoint_graph_b200.synthetic import make_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = torch.device("cuda:0")
b = make_batch(n_nodes=n, seed=5, npts=1, minpts=1)
N, E, H = b["degs"].numel(), b["idxn"].numel(), 32
graph = ops.EccGraph(b["idxn"], None, b["degs"], n_in=N)
x, g = torch.randn(N, H, device=dev), torch.randn(N, H, device=dev)
for mode in ("vv", "mat"):
    w = torch.randn((E, H, H) if mode == "mat" else (E, H), device=dev)
    gw = torch.empty_like(w)
    for _ in range(2):
        ops.ecc_fwd(x, w, graph, H)
        ops.ecc_bwd_x(w, g, graph, H)
        ops.ecc_bwd_w(x, g, graph, tuple(w.shape), out=gw)
    torch.cuda.synchronize()
    del w, gw
print("done", N, E)
