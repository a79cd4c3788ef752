"""Cycle counters of the fused chain kernel's MMA warp and epilogue warp 0 (debugging aid; never a benchmark).
Needs a library built with the instrumentation:  TEXPOSE_NVCC_EXTRA=-DTP_CHAIN_PROF python -c "from texpose_b200 import _C; _C.build(force=True)""""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, mlp_tc_bwd, ops
import texpose_b200.mlp_tc_bwd as M
lib = _C.load()
orig = lib.tp_tc_heads_backward_workspace
state = {}
def fused(cfg, sv, S, per_image, *a, **k):
    state["S"], state["B"] = S, sv.geom["shape"][0]
    return _fused(cfg, sv, S, per_image, *a, **k)
_fused = M.heads_backward_fused
# patch workspace allocation: bigger buffer, keep a handle
_empty = torch.empty
def patched_ws(S, B):
    return orig(S, B) + 32 * 148 + 2
class LibProxy:
    def __getattr__(self, n):
        if n == "tp_tc_heads_backward_workspace":
            return patched_ws
        return getattr(lib, n)
M._C = type("C", (), {"load": staticmethod(lambda: LibProxy()), "call": staticmethod(_C.call)})
last = {}
_call = _C.call
def call(name, *args):
    if name == "tp_tc_heads_backward":
        last["ws"] = args[-3]
        last["n"] = args[-2]
    return _call(name, *args)
M._C.call = staticmethod(call)
exec(open(os.path.join(os.path.dirname(__file__), "train_profile.py")).read().split("for it in range")[0])
import ctypes
for it in range(3):
    for p in g.parameters():
        p.grad = None
    ret = g.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
    var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords); var.update(ret)
    g.compute_loss(opt, var, mode="train")["all"].backward()
torch.cuda.synchronize()
S, B = 16 * 256 * 128, 16
need = orig(S, B)
off = (need + 1) & ~1
n = last["n"]
buf = (ctypes.c_float * n).from_address(0)  # placeholder
ws_ptr = last["ws"].value
t = torch.empty(32 * 148, dtype=torch.float32, device="cuda")
ctypes.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ws_ptr + off * 4), 32 * 148 * 4, 3)
c = t.view(torch.int64).view(148, 16).cpu()
tiles = c[:, 5].clamp(min=1).double()
per_stage = lambda col: (c[:, col].double() / tiles / 6).mean().item()
print("per stage cycles: total %.0f  wait weights %.0f  wait epilogue %.0f  wait h3 tile %.0f  issue %.0f" % tuple(per_stage(i) for i in range(5)))
print("epilogue warp 0 per stage: wait acc %.0f  wait h3 %.0f  wait prev store %.0f  convert %.0f  barrier %.0f" % tuple(per_stage(i) for i in range(8, 13)))
