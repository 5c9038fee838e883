"""Per-parameter gradient error of one bench-config training step against the float64 oracle, for the
default kernels, for SPG_TC=0 (exact-fp32 FMA GEMMs) and for the float32 CPU oracle itself."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superpoint_graph_b200 import ops, workloads
from superpoint_graph_b200.trainer import HostBatch, Trainer, create_model
from oracle import nets_ref

dev = torch.device("cuda:0")
w = workloads.get(sys.argv[1] if len(sys.argv) > 1 else "s3dis_train")
batch = workloads.batch(w, 1)
f64 = lambda d: {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
pcfg, mcfg = workloads.oracle_cfg(w["margs"])


def oracle(double):
    torch.manual_seed(1)
    m = create_model(w["margs"])
    cv = f64 if double else dict
    r = nets_ref.RefTrainer(cv(m.ptn.state_dict()), cv(m.ecc.state_dict()), pcfg, mcfg, ecc_mode="vec")
    r.step(f64(batch) if double else batch)
    g = {}
    for pre, sd in (("ecc.", r.sd_ecc), ("ptn.", r.sd_ptn)):
        for k, v in sd.items():
            if nets_ref.is_param(k):
                g[pre + k] = v.grad.double()
    return g


def gpu(tc):
    ops.USE_TC[0] = tc
    torch.manual_seed(1)
    m = create_model(w["margs"]).to(dev)
    tr = Trainer(m, w["margs"])
    tr.train_step(HostBatch(batch).to_device(dev))
    return {k: p.grad.double().cpu() for k, p in m.named_parameters()}


truth = oracle(True)
rows = {"cpu_f32": oracle(False), "gpu_tc": gpu(True), "gpu_fma": gpu(False)}
print("%-28s %10s | %s" % ("parameter", "max|g|", "  ".join("%9s" % k for k in rows)))
for k, t in truth.items():
    sc = max(float(t.abs().max()), 1e-30)
    print("%-28s %10.3e | %s" % (k, sc, "  ".join("%9.2e" % (float((rows[r][k] - t).abs().max()) / sc) for r in rows)))
