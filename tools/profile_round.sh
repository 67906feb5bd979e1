#!/bin/bash
# Profiles of one round on the GPU box (outputs under gpurun_out/, summarised into profiles/ afterwards):
#   the ncu launch list of a short bench run, a full ncu capture of the dominant kernel (ndt_align_kernel: one launch per
#   align since round 2 - no host command channel, so Nsight Compute can replay it) and of the smaller kernels.
TAG=${1:-prof}
set -x
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --loop-pairs 8 --loop-host-pairs 0 --odometry-sweeps 6 --no-big-map > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -c 2 -f -o gpurun_out/${TAG}_prof_ndt_align python tools/dev/dev_ndt_align_trace.py > gpurun_out/${TAG}_ncu_ndt.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"voxel_sums|voxel_stats|nn_knn|gicp_linearize|gicp_correspondence|radix_onesweep" -c 12 -f -o gpurun_out/${TAG}_prof_small python tools/bench_stages.py --reps 1 > gpurun_out/${TAG}_ncu_small.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
