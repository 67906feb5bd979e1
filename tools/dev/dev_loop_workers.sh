#!/bin/bash
# development aid (GPU box): loop-closure throughput of bench.py against the number of concurrent host workers
for w in 1 4 8; do
  LGS_BENCH_LOOP_WORKERS=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/lw.json
  python - <<PY
import json
d = json.load(open("/tmp/lw.json"))
l = d["loop_closure"]
print("workers", $w, "host arrays %.1f pairs/s" % l["pairs_per_sec"], "key-frame array", l["from_keyframe_array"].get("pairs_per_sec"))
PY
done
