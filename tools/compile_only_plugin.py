"""pytest plugin for a box WITHOUT a GPU: dry run of the gpu-marked tests in which every oracle constructor only
COMPILES its problem and then stops the test with a sentinel, to see whether any problem the GPU tests build is
rejected by the compiler after a change to the rules.

    PYTHONPATH=tools python -m pytest tests -m gpu -p compile_only_plugin -q -rf --tb=no \
        --deselect tests/test_sharded_nccl.py --deselect tests/test_gpu_fullsize_properties.py | grep -v CompiledOnly

Every test then "fails"; anything that is not ``CompiledOnly`` deserves a look (expected leftovers: the atom fixtures
whose Hessian rule the reference rejects - the dry run ignores ``with_hessian=False`` - and the C-ABI error-path test)."""
import pytest


class CompiledOnly(Exception):
    pass


@pytest.fixture(autouse=True)
def _compile_only(monkeypatch):
    from dnlp_b200 import oracles, multistart, _cabi
    from dnlp_b200.compiler import compile_problem

    def init(self, problem, *a, tape=None, with_hessian=True, **k):
        if tape is None:
            compile_problem(problem, with_hessian=with_hessian)
        raise CompiledOnly("compiled")
    monkeypatch.setattr(oracles.GpuOracles, "__init__", init)

    def binit(self, problem, *a, **k):
        compile_problem(problem)
        raise CompiledOnly("compiled")
    monkeypatch.setattr(multistart.BatchedOracles, "__init__", binit)
    monkeypatch.setattr(_cabi, "device_count", lambda: 1)
    yield


def pytest_configure(config):
    from dnlp_b200 import _cabi
    _cabi.device_count = lambda: 1          # let the gpu-marked tests run (their oracles stop at the sentinel)
