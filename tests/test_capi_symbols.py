"""The C-ABI library loads on a CPU-only box and exports every symbol include/srps_c_api.h
declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "srps_c_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srps_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import srmeetsps_cuda_b200.build as b
    lib_path = b.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in srps_c_api.h but not exported"


def test_python_binding_covers_header():
    from srmeetsps_cuda_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    lib = _lib.load()
    assert b"sm_100a" in lib.srps_build_info()


def test_create_rejects_bad_arguments_without_gpu():
    """Argument validation happens before any CUDA call."""
    import numpy as np
    from srmeetsps_cuda_b200 import Context, SRPSError
    mask = np.ones((8, 8), np.uint8)
    K = [10, 0, 0, 0, 10, 0, 4, 4, 1]
    with pytest.raises(SRPSError, match="n_channels"):
        Context(mask, 4, 2, K, n_channels=1)
    with pytest.raises(SRPSError, match="sf must be"):
        Context(mask, 4, 3, K)
    with pytest.raises(SRPSError, match="multiples of sf"):
        Context(np.ones((6, 8), np.uint8), 4, 4, K)
    with pytest.raises(SRPSError, match="n_images"):
        Context(mask, 0, 2, K)
