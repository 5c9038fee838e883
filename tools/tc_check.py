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
