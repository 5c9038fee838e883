"""Timeline probe of Trainer.eval_step_host: e2e ms for several chunk counts + where the time goes."""
import sys
import time

import torch

sys.path.insert(0, ".")
from superpoint_graph_b200 import ops, workloads  # noqa: E402
from superpoint_graph_b200.spg_pointnet import CloudEmbedder  # noqa: E402
from superpoint_graph_b200.trainer import HostBatch, Trainer, create_model  # noqa: E402

dev = torch.device("cuda", 0)
for name, nodes in (("sema3d_eval", 20000), ("vkitti_eval", 8192)):
    w = workloads.get(name, nodes)
    torch.manual_seed(1)
    model = create_model(w["margs"]).to(dev)
    tr = Trainer(model, w["margs"], dtype=w["dtype"])
    hb = HostBatch(workloads.batch(w, 1))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = torch.empty((nodes, w["margs"].classes), dtype=torch.float32).pin_memory()

    def run(fn, n=12):
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        tot, host = 0.0, 0.0
        for _ in range(n):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            t0 = time.perf_counter()
            lg = fn()
            host += time.perf_counter() - t0
            out_host.copy_(lg, non_blocking=True)
            e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
        return tot / n, host / n * 1e3

    db = hb.to_device(dev)
    print(name, "resident      gpu %.3f ms  host-issue %.3f ms" % run(lambda: tr.eval_step(db)))
    print(name, "plain upload  gpu %.3f ms  host-issue %.3f ms" % run(lambda: tr.eval_step(hb.to_device(dev))))
    # upload alone
    def up():
        hb.to_device(dev)
        return db.labels[:1].float().expand(nodes, w["margs"].classes) if False else torch.zeros(1, device=dev).expand(nodes, w["margs"].classes).contiguous()
    print(name, "upload only   gpu %.3f ms  host-issue %.3f ms" % run(up))
    for ch in (1, 2, 3, 4, 6, 8):
        CloudEmbedder.PIPELINE_CHUNKS = ch
        print(name, "pipelined x%d  gpu %.3f ms  host-issue %.3f ms" % ((ch,) + run(lambda: tr.eval_step_host(hb))))
