#!/bin/bash
# last pass of the round: smoke, default bench line, the three inference lines (with cpu_baseline and roofline)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
for spec in "room_fwd 1536 30" "sema3d_eval 20000 20" "vkitti_eval 8192 30"; do set -- $spec
timeout 600 python bench.py --workload $1 --nodes $2 --steps $3 --warmup 5 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err
done
python - <<'PY'
import json
for f in ["default", "room_fwd", "sema3d_eval", "vkitti_eval"]:
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json' % f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "FAILED", ex); print(open('gpurun_out/r2_bench_%s.err' % f).read()[-1500:]); continue
    r = d.get('roofline') or {}
    cb = d.get('cpu_baseline') or {}
    print(f, "ms/step %.3f" % d['ms_per_step'], "e2e ms %.3f" % d['e2e'].get('ms_per_step', 0), "launches", d.get('gpu_launches'),
          "roofline", r.get('kernel'), r.get('frac'), "cpu ms", cb.get('ms_per_step'), "parity", d.get('parity_rel_err'))
PY
