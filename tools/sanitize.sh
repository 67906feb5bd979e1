#!/bin/bash
# compute-sanitizer over the GPU parity suite on the box (memcheck / racecheck in the test plan).
# Skips the full-size cases (20 M-point map, 4096-pair batch, 262 144-point sweeps): under the tools a kernel runs 10-100x
# slower and those only repeat the same kernels on more data.  Outputs: gpurun_out/<tag>_{memcheck,racecheck,synccheck}.log
TAG=${1:-san}
K='velodyne or device_resident or non_finite or hash_table or always_rebuilds or outlier or radix or knn_exact or covariances or pointcloud2 or keyframe_array or icp_step or matches_single_pair or degenerate or device_resident_inputs'
for TOOL in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/${TAG}_${TOOL}.log 2>&1
  echo "$TOOL exit $?" >> gpurun_out/${TAG}_${TOOL}.log
  tail -4 gpurun_out/${TAG}_${TOOL}.log
done
