import os
import sys

import numpy as np
import pytest
import torch

# the parity tests assert the reference's fp32 numbers (<= 1e-4): pin the MLP mode for the suite; the tensor-core tests select
# bf16 explicitly through opt.b200.mlp (the library default is 'auto' = bf16 wherever the fused kernel applies)
os.environ.setdefault("TEXPOSE_B200_MLP", "fp32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    """npz fixture as torch tensors (scalars stay python numbers)."""

    def __getattr__(self, k):
        return self[k]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = Golden()
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.ndim > 0 else a.item()
    return out


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def layer_list(module_list):
    return [(l.weight.detach(), l.bias.detach()) for l in module_list]


def check_weight_checksums(module, g):
    """The fixtures carry checksums of the reference's seed-0 weights; the product module must re-create them."""
    n = 0
    for name, p in module.named_parameters():
        if p.dim() == 0:
            continue
        assert abs(p.detach().double().sum().item() - g["ck_sum/" + name]) < 1e-9, name
        assert abs(p.detach().double().abs().sum().item() - g["ck_abs/" + name]) < 1e-9, name
        assert torch.equal(p.detach().flatten()[:4].cpu(), g["ck_head/" + name]), name
        n += 1
    assert n > 0
