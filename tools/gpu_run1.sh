#!/bin/bash
# round 2, first GPU pass: parity at the benchmarked shapes + bench lines of every workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -30 gpurun_out/r2_pytest1.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_s3dis.json 2> gpurun_out/r2_bench_s3dis.err; tail -c 1500 gpurun_out/r2_bench_s3dis.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json
for w in room_fwd sema3d_eval; do
python bench.py --workload $w --steps 20 --warmup 5 --no-roofline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; cat gpurun_out/r2_bench_$w.json; tail -3 gpurun_out/r2_bench_$w.err
done
python bench.py --workload sweep_vv --nodes 10000 --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench_sweep10k.json 2> gpurun_out/r2_bench_sweep10k.err; cat gpurun_out/r2_bench_sweep10k.json; tail -3 gpurun_out/r2_bench_sweep10k.err
