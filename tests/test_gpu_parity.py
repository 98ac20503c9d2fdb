"""GPU parity: the CUDA tape through the C-ABI vs (a) the live-reference golden vectors and
(b) the CPU oracle on fresh seeded points.  Structures bit-exact, values rel 1e-10 (fp64)."""
import numpy as np
import pytest

from golden_util import Golden, assert_close, golden_names
from oracle.dnlp_oracle import RefOracles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_mod():
    from dnlp_b200.oracles import GpuOracles
    return GpuOracles


@pytest.mark.parametrize("name", golden_names())
def test_gpu_matches_reference_golden(name, gpu_mod):
    g = Golden(name)
    o = gpu_mod(g.problem)
    try:
        jr, jc = o.jacobianstructure()
        hr, hc = o.hessianstructure()
        assert jr.dtype == np.int32 and hr.dtype == np.int32
        np.testing.assert_array_equal(jr, g.jac_rows)
        np.testing.assert_array_equal(jc, g.jac_cols)
        np.testing.assert_array_equal(hr, g.hess_rows)
        np.testing.assert_array_equal(hc, g.hess_cols)
        for i, p in enumerate(g.points):
            assert_close(o.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i)
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
        # interleaved call order with the x-keyed cache active (IPOPT's pattern: same x, five calls)
        for p in reversed(g.points):
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess/cached")
            assert_close(o.jacobian(p["x"]), p["jac"], "jac/cached")
            assert_close(o.objective(p["x"]), p["f"], "f/cached")
        p = g.points[-1]
        res = o.eval_all(p["x"], p["lam"], float(p["sigma"]))
        for k in ("f", "grad", "g", "jac", "hess"):
            assert_close(res[k], p[k], "eval_all/" + k)
    finally:
        o.close()


@pytest.mark.parametrize("name", ["c3_logistic_small", "c5_microbench_small", "matmul_with_atom_operand",
                                  "clnlbeam", "rel_entr_vector", "hyperbolic_mix"])
def test_gpu_matches_oracle_on_fresh_points(name, gpu_mod):
    g = Golden(name)
    ref = RefOracles(g.problem)
    ref.jacobianstructure(), ref.hessianstructure()
    o = gpu_mod(g.problem)
    rng = np.random.default_rng(1234)
    try:
        x0 = g.points[0]["x"]
        with np.errstate(all="ignore"):
            for _ in range(5):
                x = x0 * (1 + 0.02 * rng.standard_normal(x0.size)) + 0.01 * rng.standard_normal(x0.size)
                lam = rng.standard_normal(g.problem.m)
                sigma = float(rng.uniform(0.1, 2.0))
                assert_close(o.objective(x), ref.objective(x), "f")
                assert_close(o.gradient(x), ref.gradient(x), "grad")
                assert_close(o.constraints(x), ref.constraints(x), "g")
                assert_close(o.jacobian(x), ref.jacobian(x), "jac")
                assert_close(o.hessian(x, lam, sigma), ref.hessian(x, lam, sigma), "hess")
    finally:
        o.close()


@pytest.mark.parametrize("name", ["c2_eigen_qcqp_small", "c3_logistic_small", "c4_qcqp_small",
                                  "c5_microbench_small", "portfolio_socp", "matmul_var_var"])
def test_gpu_optimised_emission_paths(name, gpu_mod, monkeypatch):
    """Large-problem kernels (two-stage reductions, SCALE, one-term POLY, scatter-accumulate) forced on."""
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "LAYER_MIN", 2)
    monkeypatch.setattr(Builder, "LONG_ROW", 3)
    monkeypatch.setattr(Builder, "CHUNK", 2)
    g = Golden(name)
    o = gpu_mod(g.problem)
    try:
        for i, p in enumerate(g.points):
            assert_close(o.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i)
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
    finally:
        o.close()


def _cmp_with_oracle(prob, gpu_mod, npoints=2, seed=0):
    ref = RefOracles(prob)
    jr, jc = ref.jacobianstructure()
    hr, hc = ref.hessianstructure()
    o = gpu_mod(prob)
    try:
        np.testing.assert_array_equal(o.jacobianstructure()[0], jr)
        np.testing.assert_array_equal(o.jacobianstructure()[1], jc)
        np.testing.assert_array_equal(o.hessianstructure()[0], hr)
        np.testing.assert_array_equal(o.hessianstructure()[1], hc)
        rng = np.random.default_rng(seed)
        with np.errstate(all="ignore"):
            for _ in range(npoints):
                x = prob.x0 * (1 + 0.01 * rng.standard_normal(prob.n))
                lam = rng.standard_normal(prob.m)
                sigma = float(rng.uniform(0.5, 1.5))
                assert_close(o.objective(x), ref.objective(x), "f")
                assert_close(o.gradient(x), ref.gradient(x), "grad")
                assert_close(o.constraints(x), ref.constraints(x), "g", atol=1e-9)
                assert_close(o.jacobian(x), ref.jacobian(x), "jac")
                assert_close(o.hessian(x, lam, sigma), ref.hessian(x, lam, sigma), "hess")
                res = o.eval_all(x, lam, sigma)
                assert_close(res["hess"], ref.hessian(x, lam, sigma), "eval_all/hess")
                assert_close(res["g"], ref.constraints(x), "eval_all/g", atol=1e-9)
    finally:
        o.close()


@pytest.mark.parametrize("n", [1500, 1501])
def test_gpu_medium_eigen_qcqp(n, gpu_mod):
    """Sizes that reach the production kernels: CTA-per-row GEMV (even n) / warp-per-row GEMV
    (odd n), SCALE + scatter-accumulate Hessian, grid-wide single-row reduction."""
    from dnlp_b200 import workloads as W
    _cmp_with_oracle(W.eigen_qcqp(n), gpu_mod)


def test_gpu_medium_logistic(gpu_mod):
    from dnlp_b200 import workloads as W
    At, x0 = W.logistic_data(30000, 64, 16)
    _cmp_with_oracle(W.logistic_regression(At, x0), gpu_mod)


def test_gpu_medium_microbench(gpu_mod):
    from dnlp_b200 import workloads as W
    A, x0 = W.microbench_data(80000, 40000, 10)
    _cmp_with_oracle(W.microbench(A, x0), gpu_mod)


@pytest.mark.parametrize("name", ["c3_logistic_small", "clnlbeam", "portfolio_socp", "hs071", "c2_eigen_qcqp_small"])
def test_gpu_constant_entry_elision(name, gpu_mod, monkeypatch):
    """Compact D2H of only the x/lambda-dependent entries (affine rows cached, reference quirk Q5)."""
    monkeypatch.setattr(gpu_mod, "ELIDE_MIN", 1)
    monkeypatch.setattr(gpu_mod, "ELIDE_MAX_FRACTION", 1.0)
    g = Golden(name)
    o = gpu_mod(g.problem)
    try:
        assert o._dyn, "elision not active"
        for i, p in enumerate(g.points):
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i)
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
    finally:
        o.close()
