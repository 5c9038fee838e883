#!/bin/bash
# ncu --set full of spg_cloud_build at the loader-roofline size (bench.loader_roofline), one launch
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cloud_build -s 3 -c 1 -f -o gpurun_out/r2_cloud_build \
    python -c "import bench, torch; print(bench.loader_roofline(torch.device('cuda', 0), bench.peaks(), False))" > gpurun_out/prof_loader.log 2>&1
tail -3 gpurun_out/prof_loader.log
python __graft_entry__.py smoke 2>&1 | tail -2
ls -la gpurun_out/r2_cloud_build.ncu-rep
