#!/bin/bash
# Same-box A/B of two builds of the library through bench.py (plain runs).  usage: bash scripts/ab_lib.sh /path/to/other.so
for i in 1 2 3; do
  for lib in "" "$1"; do
    TEXPOSE_B200_LIB=$lib timeout -s KILL 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib', '${lib:-default}'[-16:], 'ms/frame %.2f' % d['ms_per_step'], 'kernel_ms %.2f' % d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
  done
done
