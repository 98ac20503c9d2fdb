"""The C-ABI library loads and exports every symbol include/dnlp_b200.h declares (no compute)."""
import ctypes
import os
import re

from dnlp_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dnlp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dnlp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_cabi.EXPORTS) == names


def test_no_silent_cpu_fallback():
    """Without a CUDA device creating an oracle must raise, never compute on the host."""
    import numpy as np
    import pytest
    from dnlp_b200 import ir
    from dnlp_b200.oracles import GpuOracles
    if _cabi.device_count() > 0:
        pytest.skip("CUDA device present")
    x = ir.Variable(3)
    prob = ir.ProblemIR(ir.sum(ir.exp(x)), [], x0=np.zeros(3))
    with pytest.raises(RuntimeError):
        GpuOracles(prob)
