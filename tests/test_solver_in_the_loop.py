"""Solver-in-the-loop check of the drop-in boundary.

IPOPT/cyipopt is not installed in the build container or on the GPU boxes, so a small
equality-constrained Newton-KKT solver (tests/kkt_newton.py) stands in for it: it touches the oracle
only through the cyipopt callback surface.  On the README toy (BASELINE config 1) it must reach the
published optimum lambda_max(A) = 11.950810853979528 (reference README.md:26-54) with the CPU oracle,
and the GPU oracle must reproduce the same iteration count and optimum within 1e-8."""
import numpy as np
import pytest

import kkt_newton
from dnlp_b200 import workloads as W
from oracle.dnlp_oracle import RefOracles

README_OPTIMUM = 11.950810853979528


def test_readme_toy_reaches_published_optimum_with_cpu_oracle():
    p = W.eigen_qcqp(3)
    x, lam, f, iters = kkt_newton.solve(RefOracles(p), p.x0)
    assert abs(-f - README_OPTIMUM) < 1e-8
    assert abs(np.sum(x ** 2) - 1.0) < 1e-10
    assert iters < 30


@pytest.mark.gpu
@pytest.mark.parametrize("n", [3, 24, 200])
def test_gpu_oracle_reproduces_iteration_count_and_optimum(n):
    from dnlp_b200.oracles import GpuOracles
    p = W.eigen_qcqp(n)
    xr, lr, fr, itr = kkt_newton.solve(RefOracles(p), p.x0)
    o = GpuOracles(p)
    try:
        xg, lg, fg, itg = kkt_newton.solve(o, p.x0)
    finally:
        o.close()
    assert itg == itr
    assert abs(fg - fr) <= 1e-8 * max(1.0, abs(fr))
    np.testing.assert_allclose(xg, xr, rtol=1e-8, atol=1e-10)
    if n == 3:
        assert abs(-fg - README_OPTIMUM) < 1e-8
