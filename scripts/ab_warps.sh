for i in 1 2 3; do for f in 0 2; do
  TEXPOSE_TC_FLAGS=$f timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flags $f', 'ms/frame %.2f' % d['ms_per_step'], 'kernel_ms %.2f' % d['roofline']['kernel_ms'], 'e2e %.2f' % d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done; done
