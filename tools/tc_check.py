"""Device check of the tcgen05 3xTF32 GEMM against fp64 (run under `timeout`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from superpoint_graph_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True
for (M, N, K) in [(128, 64, 32), (128, 64, 64), (1000, 64, 64), (4096, 128, 64), (5000, 128, 128), (3000, 256, 128),
                  (130, 64, 256), (120576, 256, 128)]:
    A = torch.randn(M, K, dtype=torch.float64)
    W = torch.randn(N, K, dtype=torch.float64)
    b = torch.randn(N, dtype=torch.float64)
    sc, sh = torch.rand(K, dtype=torch.float64) + 0.5, torch.randn(K, dtype=torch.float64)
    Af, Wf = A.float().to(dev), W.float().to(dev)
    ref = Af.double() @ Wf.double().t() + b.to(dev)
    out, mean, var = ops.tc_gemm(Af, K, Wf, K, False, M, N, K, bias=b.float().to(dev), stats=True)
    torch.cuda.synchronize()
    e1 = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    em = ((mean.double() - ref.mean(0)).abs().max() / ref.mean(0).abs().max()).item()
    ev = ((var.double() - ref.var(0, unbiased=False)).abs().max() / ref.var(0, unbiased=False).abs().max()).item()
    ref2 = torch.relu(Af.double() * sc.to(dev) + sh.to(dev)) @ Wf.double().t()
    out2 = ops.tc_gemm(Af, K, Wf, K, False, M, N, K, a_aff=(sc.float().to(dev), sh.float().to(dev), True))
    e2 = ((out2.double() - ref2).abs().max() / ref2.abs().max()).item()
    # transposed weights (data gradient): C[M,K] = dY[M,N] W[N,K]
    if K in (64, 128, 256) and N % 32 == 0:
        dY = torch.randn(M, N, device=dev)
        ref3 = dY.double() @ Wf.double()
        out3 = ops.tc_gemm(dY, N, Wf, K, True, M, K, N)
        e3 = ((out3.double() - ref3).abs().max() / ref3.abs().max()).item()
    else:
        e3 = 0.0
    good = max(e1, e2, e3) < 2e-5 and em < 1e-4 and ev < 1e-4
    ok &= good
    print("M=%d N=%d K=%d  err %.2e  prologue %.2e  transposed %.2e  mean %.2e var %.2e  %s"
          % (M, N, K, e1, e2, e3, em, ev, "ok" if good else "FAIL"), flush=True)
# fused BatchNorm modes: statistics + fold, BN-backward prologue (+ dY side store), BN-backward sums
for (M, N, K) in [(1000, 64, 64), (5000, 128, 128), (120576, 128, 256), (120576, 64, 64), (9654, 32, 32), (777, 128, 64)]:
    # forward with statistics + fold
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.3
    gamma, beta = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
    rm, rv, nbt = torch.zeros(N, device=dev), torch.ones(N, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    y, mean, var, scale, shift = ops.tc_gemm(A, K, W, K, False, M, N, K, stats=True,
                                              fold=(gamma, beta, 1e-5, rm, rv, nbt, 0.1))
    yd = A.double() @ W.double().t()
    mu_r, var_r = yd.mean(0), yd.var(0, unbiased=False)
    sc_r = gamma.double() / torch.sqrt(var_r + 1e-5)
    e_m = ((mean.double() - mu_r).abs().max() / mu_r.abs().max()).item()
    e_v = ((var.double() - var_r).abs().max() / var_r.abs().max()).item()
    e_s = ((scale.double() - sc_r).abs().max() / sc_r.abs().max()).item()
    e_t = ((shift.double() - (beta.double() - mu_r * sc_r)).abs().max() / (beta.double() - mu_r * sc_r).abs().max()).item()
    e_rv = ((rv.double() - (0.9 + 0.1 * var_r * M / (M - 1))).abs().max()).item()
    # backward: G = dL/d(relu(bn(y))) ; dX = dY @ Wd with dY = BN/ReLU backward of G, Wd [N, K2]
    K2 = 64 if N != 32 else 32
    Wd = torch.randn(N, K2, device=dev) * 0.3        # the layer's weight [cout=N, cin=K2]
    G = torch.randn(M, N, device=dev)
    yr = y.double().clone().requires_grad_(True)  # the kernel's own y: ReLU masks of near-zero elements must agree
    xhat = (yr - yr.mean(0)) / torch.sqrt(yr.var(0, unbiased=False) + 1e-5)  # batch statistics inside the graph
    act = torch.relu(xhat * gamma.double() + beta.double())
    act.backward(G.double())
    dY_r = yr.grad
    gz = G.double() * (act > 0)
    s1_r, s2_r = gz.sum(0), (gz * xhat.detach()).sum(0)
    s12 = torch.cat([s1_r, s2_r]).float()
    # the layer below (for the epilogue sums): raw y2 [M,K2] with its own fold
    y2 = torch.randn(M, K2, device=dev)
    m2_, v2_ = y2.double().mean(0), y2.double().var(0, unbiased=False)
    g2, b2 = torch.rand(K2, device=dev) + 0.5, torch.randn(K2, device=dev) * 0.2
    sc2 = (g2.double() / torch.sqrt(v2_ + 1e-5)).float()
    sh2 = (b2.double() - m2_ * sc2.double()).float()
    dX, dY, red = ops.tc_gemm(G, N, Wd, K2, True, M, K2, N,
                              bnbwd=(y, N, scale, shift, True, mean, var, s12, 1e-5, True),
                              bnred=(y2, K2, sc2, sh2, m2_.float(), v2_.float(), 1e-5, True))
    torch.cuda.synchronize()
    dX_r = dY_r @ Wd.double()
    e_dy = ((dY.double() - dY_r).abs().max() / dY_r.abs().max()).item()
    e_dx = ((dX.double() - dX_r).abs().max() / dX_r.abs().max()).item()
    xh2 = (y2.double() - m2_) / torch.sqrt(v2_ + 1e-5)
    gz2 = dX_r * ((xh2 * g2.double() + b2.double()) > 0)
    r1, r2 = gz2.sum(0), (gz2 * xh2).sum(0)
    e_r = max(((red[:K2].double() - r1).abs().max() / r1.abs().max()).item(),
              ((red[K2:].double() - r2).abs().max() / r2.abs().max()).item())
    good = max(e_m, e_v, e_s, e_t) < 1e-4 and e_rv < 1e-4 and e_dy < 1e-4 and e_dx < 1e-4 and e_r < 2e-3
    ok &= good
    print("fused M=%d N=%d K=%d: mean %.1e var %.1e scale %.1e shift %.1e rvar %.1e | dY %.1e dX %.1e s12 %.1e  %s"
          % (M, N, K, e_m, e_v, e_s, e_t, e_rv, e_dy, e_dx, e_r, "ok" if good else "FAIL"), flush=True)
for (M, co, ci) in [(64, 128, 64), (4096, 128, 64), (5000, 128, 128), (100000, 256, 128), (3333, 256, 64)]:
    dY = torch.randn(M, co, device=dev)
    P = torch.randn(M, ci, device=dev)
    sc, sh = torch.rand(ci, device=dev) + 0.5, torch.randn(ci, device=dev)
    ref = dY.double().t() @ torch.relu(P.double() * sc.double() + sh.double())
    out = ops.tc_dw(dY, co, P, ci, M, co, ci, p_aff=(sc, sh, True))
    torch.cuda.synchronize()
    e = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    ref0 = dY.double().t() @ P.double()
    e0 = ((ops.tc_dw(dY, co, P, ci, M, co, ci).double() - ref0).abs().max() / ref0.abs().max()).item()
    good = max(e, e0) < 2e-5
    ok &= good
    print("dW M=%d co=%d ci=%d  err %.2e (plain %.2e) %s" % (M, co, ci, e, e0, "ok" if good else "FAIL"), flush=True)
print("ALL OK" if ok else "SOME FAILED")
