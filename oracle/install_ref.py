"""Install the reference for the bench's reference arm:  python -m oracle.install_ref

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (HanzhiC/TexPose) is a plain Python checkout with no setup.py / pyproject,
so `pip install /root/reference` does not apply; its .py and .yaml files (540 KB; `external/` and `splits/` are not on the
path) are copied, unmodified, into baseline/_ref -- git-ignored, NOT gpurun-ignored, so the copy travels to the GPU box
where /root/reference does not exist.  `bench.py --impl reference`, the eager-GPU "before" number and the Option-B
integration test import it from there through oracle/ref_import.py.  Nothing under texpose_b200/ reads it.
"""
from __future__ import annotations

import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
SKIP_DIRS = {"external", "splits", ".git", "__pycache__"}


def install(src: str = SRC, dst: str = DST) -> int:
    """Copies the reference's sources; returns the number of files (0 when the checkout is absent)."""
    if not os.path.isdir(os.path.join(src, "layers")):
        return 0
    n, digest = 0, hashlib.sha256()
    for base, dirs, files in os.walk(src):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if not f.endswith((".py", ".yaml", ".md")) and f != "LICENSE":
                continue
            rel = os.path.relpath(os.path.join(base, f), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), out)
            digest.update(rel.encode())
            digest.update(open(out, "rb").read())
            n += 1
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as fh:
        fh.write(f"{src}\nfiles {n}\nsha256 {digest.hexdigest()}\nunmodified copy made by oracle/install_ref.py\n")
    return n


if __name__ == "__main__":
    print(f"installed {install()} reference files into {DST}")
