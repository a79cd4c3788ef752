"""Cycle counters of the fused chain kernel's MMA warp and epilogue warp 0 (debugging aid; never a benchmark).

Needs a library built with the instrumentation:
    TEXPOSE_NVCC_EXTRA=-DTP_CHAIN_PROF python -c "from texpose_b200 import _C; _C.build(force=True)"
    TEXPOSE_CHAIN_PROF=1 python scripts/chain_prof.py
(rebuild without the define afterwards: the counters cost a few percent)."""
import os
import sys

import torch

os.environ["TEXPOSE_CHAIN_PROF"] = "1"
sys.argv = [sys.argv[0], "3"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_profile.py")).read())
from texpose_b200 import mlp_tc_bwd  # noqa: E402

ws, off = mlp_tc_bwd.last_profile_workspace
c = ws[off:off + 32 * 148].view(torch.int64).view(148, 16).cpu()
tiles = c[:, 5].clamp(min=1).double()
per_stage = lambda col: (c[:, col].double() / tiles / 6).mean().item()
print("MMA warp, cycles per stage: total %.0f  wait weights %.0f  wait epilogue %.0f  wait h3 tile %.0f  issue %.0f"
      % tuple(per_stage(i) for i in range(5)))
print("epilogue warp 0, cycles per stage: wait acc %.0f  wait h3 %.0f  wait prev store %.0f  convert %.0f  barrier %.0f"
      % tuple(per_stage(i) for i in range(8, 13)))
