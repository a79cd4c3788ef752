"""The C-ABI library loads on a CPU-only box and exports every symbol include/texpose_b200.h declares."""
import ctypes
import os

from texpose_b200 import _C


def test_header_symbols_exported():
    _C.build()
    protos = _C.declared_prototypes()
    names = [p[1] for p in protos]
    assert len(names) >= 25 and len(set(names)) == len(names)
    lib = ctypes.CDLL(_C.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_version_and_prototypes():
    lib = _C.load()
    assert lib.tp_version() == 100
    for ret, name, args in _C.declared_prototypes():
        fn = getattr(lib, name)
        assert len(fn.argtypes) == len(args)


def test_no_torch_types_in_abi():
    import re
    text = re.sub(r"/\*.*?\*/", " ", open(_C.HEADER).read(), flags=re.S)     # declarations only
    assert "torch" not in text and "at::" not in text and "Tensor" not in text


def test_sources_are_sm100a_only():
    assert "arch=compute_100a,code=sm_100a" in " ".join(_C.NVCC_FLAGS)
    assert os.path.exists(_C.LIB_PATH)
