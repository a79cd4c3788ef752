#!/bin/bash
# A/B of kernel variants through bench.py (plain runs).  usage: bash scripts/ab_bench.sh "0 2"
for f in $1; do
  TEXPOSE_TC_FLAGS=$f timeout -s KILL 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | tail -1 > /tmp/ab_$f.json
  python - "$f" <<'PY'
import json, sys
d = json.load(open(f"/tmp/ab_{sys.argv[1]}.json"))
print("flags", sys.argv[1], "ms/frame %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "kernel_ms %.2f" % d["roofline"]["kernel_ms"], d["clocks"])
PY
done
