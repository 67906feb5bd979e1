#!/bin/bash
# Profiles of one round on the GPU box (outputs under gpurun_out/, summarised into profiles/ afterwards):
#   stage timings, the ncu launch list of a short bench run, full ncu captures of the derivative kernel and of the
#   smaller kernels.  The persistent evaluators are switched off for the ncu passes (Nsight Compute serialises kernels
#   and blocks the host inside the launch call; they would also switch themselves off, see persist.cuh).
TAG=${1:-prof}
set -x
timeout 500 python tools/bench_stages.py --reps 10 > gpurun_out/${TAG}_stages.json 2> gpurun_out/${TAG}_stages.err
tail -3 gpurun_out/${TAG}_stages.err
export LGS_NDT_PERSISTENT=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --loop-pairs 2 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ndt_derivatives_kernel -c 6 -f -o gpurun_out/${TAG}_prof_ndt python tests/diag_ndt_deriv.py --no-oracle --reps 2 > gpurun_out/${TAG}_ncu_ndt.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pgicp_functor|icp_step|pgicp_mahal|pgicp_corr|gicp_linearize" -c 10 -f -o gpurun_out/${TAG}_prof_small python tools/bench_stages.py --reps 1 > gpurun_out/${TAG}_ncu_small.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
