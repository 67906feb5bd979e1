set -x
timeout 400 python tools/bench_stages.py --reps 10 > gpurun_out/s8_stages.json 2> gpurun_out/s8_stages.err
tail -3 gpurun_out/s8_stages.err
export LGS_NDT_PERSISTENT=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/s8_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --loop-pairs 2 > gpurun_out/s8_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ndt_derivatives_kernel -c 6 -f -o gpurun_out/s8_prof_ndt python tests/diag_ndt_deriv.py --no-oracle --reps 2 > gpurun_out/s8_ncu_ndt.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pgicp_functor|icp_step|pc2_repack|submap_assemble|pgicp_mahal|pgicp_corr" -c 12 -f -o gpurun_out/s8_prof_new python tools/bench_stages.py --reps 1 > gpurun_out/s8_ncu_new.log 2>&1
ls -la gpurun_out/*.ncu-rep
