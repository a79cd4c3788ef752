"""C2-size ray-bias table and Philox depth kernels, event-timed.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C  # noqa: E402

DEV = "cuda:0"
names = ("tp_tc_ray_bias", "tp_sample_depth", "tp_composite_stl_forward", "tp_raygen", "tp_gather_rows")
times = {k: [] for k in names}
orig = _C.call


def timing_call(name, *a):
    if name in times:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        times[name].append((e0, e1))
    else:
        orig(name, *a)


_C.call = timing_call
sys.argv = [sys.argv[0], "6"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_timeline.py")).read().split("# ---- C3 training step")[0])
torch.cuda.synchronize()
for k, v in times.items():
    ms = sorted(a.elapsed_time(b) for a, b in v[3:])
    if ms:
        print(f"{k:28s} median {ms[len(ms) // 2] * 1e3:7.1f} us   best {ms[0] * 1e3:7.1f} us   ({len(ms)} launches)")
