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


def _have_cyipopt():
    try:
        import cyipopt  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.gpu
@pytest.mark.skipif(not _have_cyipopt(), reason="cyipopt / libipopt not installed (not in this image, not on the GPU box, "
                    "not installable offline): solver-level parity unverified")
def test_ipopt_same_iteration_count_and_optimum_with_and_without_the_gpu_oracle():
    """BASELINE north_star: the toy config must reproduce the same IPOPT iteration count and an optimal value
    within 1e-8.  `prob.solve(nlp=True, solver=cp.IPOPT)` (reference call path: ipopt_nlpif.py:143-170) on the
    README toy (README.md:26-54) with the reference's own Oracles, then with install(); runs wherever cyipopt
    and the shipped reference copy (oracle/_ref) are importable."""
    from oracle import ref_driver as R
    if not R.available():
        pytest.skip("oracle/_ref not shipped")
    cp = R.load_reference()
    import dnlp_b200.nlp_solver as gpu

    def solve():
        np.random.seed(0)
        n = 3
        A = np.random.randn(n, n)
        A = A.T @ A
        x = cp.Variable(n)
        x.value = np.ones(n)
        prob = cp.Problem(cp.Maximize(cp.quad_form(x, A)), [cp.sum_squares(x) == 1])
        prob.solve(nlp=True, solver=cp.IPOPT)
        return prob.value, prob.solver_stats.num_iters, np.array(x.value)

    v_ref, it_ref, x_ref = solve()
    with gpu.gpu_oracle():
        v_gpu, it_gpu, x_gpu = solve()
    assert abs(v_ref - README_OPTIMUM) < 1e-6
    assert it_gpu == it_ref
    assert abs(v_gpu - v_ref) < 1e-8
    np.testing.assert_allclose(np.abs(x_gpu), np.abs(x_ref), atol=1e-7)
