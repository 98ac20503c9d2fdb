"""pytest plugin for a box WITHOUT a GPU: dry run of gpu-marked tests on the interpreter-backed stand-in device of the
CPU tier (tests/host_logic_device.py, test infrastructure), to check a new GPU test's own harness - fixture loading,
call sequences, return conventions, tolerances - before it meets a B200.  Nothing CUDA runs; the kernels are not under
test here.

    PYTHONPATH=tools:tests python -m pytest tests/test_zz_gpu_reference_suite.py -m gpu -p standin_device_plugin -q
"""
import pytest


@pytest.fixture(autouse=True)
def _standin_device(monkeypatch):
    import host_logic_device
    host_logic_device.install(monkeypatch)
    yield


def pytest_configure(config):
    from dnlp_b200 import _cabi
    _cabi.device_count = lambda: 1          # let the gpu-marked tests run (their oracles live on the stand-in)
