"""Locates errors of the BNBWD prologue: per row-tile / per column-chunk error map of the dY side store."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, K2) in [(5000, 128, 64), (20000, 128, 64), (120576, 128, 64), (120576, 128, 128), (120576, 256, 128), (120576, 64, 64)]:
    y = torch.randn(M, N, device=dev)
    G = torch.randn(M, N, device=dev)
    gamma, beta = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
    mean, var = y.double().mean(0), y.double().var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.double() * rstd).float()
    shift = (beta.double() - mean * gamma.double() * rstd).float()
    xhat = (y.double() - mean) * rstd
    mask = (xhat * gamma.double() + beta.double()) > 0
    gz = G.double() * mask
    s1, s2 = gz.sum(0), (gz * xhat).sum(0)
    dY_r = scale.double() * (gz - s1 / M - xhat * s2 / M)
    s12 = torch.cat([s1, s2]).float()
    Wd = torch.randn(N, K2, device=dev) * 0.3
    dX, dY = ops.tc_gemm(G, N, Wd, K2, True, M, K2, N,
                         bnbwd=(y, N, scale, shift, True, mean.float(), var.float(), s12, 1e-5, True))
    torch.cuda.synchronize()
    err = (dY.double() - dY_r).abs()
    sc = dY_r.abs().max()
    tiles = (M + 127) // 128
    pad = tiles * 128 - M
    e2 = torch.nn.functional.pad(err, (0, 0, 0, pad)).view(tiles, 128, N // 32, 32).amax(dim=(1, 3)) / sc  # [tiles, chunks]
    bad = (e2 > 1e-4).nonzero()
    dXr = dY_r @ Wd.double()
    print("M=%d N=%d K2=%d: dY max rel %.2e, dX %.2e, bad (tile,chunk) cells: %d of %d" % (
        M, N, K2, float(e2.max()), float((dX.double() - dXr).abs().max() / dXr.abs().max()), bad.shape[0], e2.numel()), flush=True)
    if bad.shape[0]:
        print("   first bad cells:", bad[:12].tolist(), " last:", bad[-4:].tolist())
        t0, c0 = bad[0].tolist()
        blk = err[t0 * 128:(t0 + 1) * 128, c0 * 32:(c0 + 1) * 32] / sc
        print("   rows with errors in that cell:", (blk.amax(1) > 1e-4).nonzero().flatten()[:40].tolist())
        print("   cols with errors in that cell:", (blk.amax(0) > 1e-4).nonzero().flatten().tolist())
