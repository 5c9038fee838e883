#!/bin/bash
# round-2 final pass on one GPU: full GPU test suite, smoke, every workload's bench line, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest_final.log | head -20
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -4 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
for spec in "room_fwd 1536 30" "sema3d_eval 20000 20" "vkitti_train 1024 30" "vkitti_eval 8192 20" "sweep_vv 10000 10" "sweep_mat 10000 10"; do set -- $spec
timeout 900 python bench.py --workload $1 --nodes $2 --steps $3 --warmup 5 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err
done
python bench.py --workload room_fwd --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_room_fwd_reference.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "FAILED", ex); continue
    cb = d.get('cpu_baseline') or {}
    print(f.split('/')[-1], d.get('impl','b200'), "ms/step %.3f" % d['ms_per_step'], "value %.4g" % d['value'], "e2e ms", round(d['e2e'].get('ms_per_step', 0),3),
          "launches", d.get('gpu_launches'), "parity", d.get('parity_rel_err'), "cpu ms", cb.get('ms_per_step'), cb.get('kind'), cb.get('cores'))
PY
