#!/bin/bash
# 1 / 2 / 4 / 8-GPU evidence on one box: bench.py at every N (loop closure = 4096 pairs in total, strong scaling; NDT = one
# sequence per GPU) and the real-rank bitwise test of tests/test_gpu_dist.py.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-scale}
NMAX=${2:-8}
PORT=29700
for N in 1 2 4 8; do
  [ $N -gt $NMAX ] && break
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 40 --warmup 3 --no-cpu-baseline --odometry-sweeps 0 --no-big-map --loop-host-pairs 64 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+N)) bench.py --gpus $N --steps 40 --warmup 3 --loop-host-pairs 64 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    l=d["loop_closure"]
    print("N=${N}: NDT %.1f aligns/s | loop closure %.1f pairs/s (%d pairs, %d ranks, gather %.2f ms) | host arrays %s" % (d["value"], l["pairs_per_sec"], l["n_pairs"], l["comm_nranks"], l["gather_ms_rank0"], l.get("from_host_arrays",{}).get("pairs_per_sec")))
except Exception as e:
    print("N=${N}: failed", e)
PY
done
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${TAG}_dist_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_dist_pytest.log
