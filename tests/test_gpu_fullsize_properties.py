"""The BASELINE.json sizes themselves (C2 n = 8192, C3 m = 2 M / n = 4096, C5 N = 10 M / 50 M nnz,
C4 B = 4096): this is where int32 positions, 64-bit row pointers, chunk tables and batch offsets can
break, so every config is checked at the size the benchmark quotes.

1. Parity with the CPU oracle (oracle/dnlp_oracle.py, itself pinned to the live reference by the small
   and medium fixtures): structures bit-exact, all five outputs within rel 1e-10 at one seeded point.
2. Size-independent properties: derivative consistency by central differences along random directions
   - the check IPOPT's own `derivative_test` option performs and to which the reference's test-suite
   delegates derivative correctness - linearity of the Hessian of the Lagrangian in (sigma, lambda),
   agreement of the fused eval_all with the individual callbacks.

Full size is the default; DNLP_FULLSIZE=0 scales C3 / C5 down for quick development runs."""
import os

import numpy as np
import pytest

from golden_util import assert_close

pytestmark = pytest.mark.gpu
FULL = os.environ.get("DNLP_FULLSIZE", "1") != "0"


def _check_against_port(prob, o, seed=0):
    """GpuOracles vs the CPU oracle at one point; returns nothing, raises on any mismatch."""
    from oracle.dnlp_oracle import RefOracles
    r = RefOracles(prob)
    jr, jc = r.jacobianstructure()
    hr, hc = r.hessianstructure()
    gjr, gjc = o.jacobianstructure()
    ghr, ghc = o.hessianstructure()
    assert gjr.dtype == np.int32 and ghr.dtype == np.int32
    np.testing.assert_array_equal(gjr, jr)
    np.testing.assert_array_equal(gjc, jc)
    np.testing.assert_array_equal(ghr, hr)
    np.testing.assert_array_equal(ghc, hc)
    rng = np.random.default_rng(100 + seed)
    x = np.asarray(prob.x0, dtype=np.float64) * (1 + 0.01 * rng.standard_normal(prob.n))
    lam = rng.standard_normal(prob.m)
    sigma = 0.9
    with np.errstate(all="ignore"):
        assert_close(o.objective(x), r.objective(x), "f")
        assert_close(o.gradient(x), r.gradient(x), "grad")
        # g = t - A~x style rows cancel to ~1e-16 * |terms|: absolute floor next to rel 1e-10
        assert_close(o.constraints(x), r.constraints(x), "g", atol=1e-9)
        assert_close(o.jacobian(x), np.asarray(r.jacobian(x)).ravel(), "jac")
        assert_close(o.hessian(x, lam, sigma), np.asarray(r.hessian(x, lam, sigma)).ravel(), "hess")


def _spmv(rows, cols, vals, d, nrows):
    return np.bincount(rows, weights=vals * d[cols], minlength=nrows)


def _check_derivatives(prob, seed=0, h_scale=1e-6, rtol=2e-6, port=True):
    from dnlp_b200.oracles import GpuOracles
    o = GpuOracles(prob)
    try:
        if port:
            _check_against_port(prob, o, seed)
        rng = np.random.default_rng(seed)
        n, m = prob.n, prob.m
        x = np.asarray(prob.x0, dtype=np.float64) * (1 + 0.01 * rng.standard_normal(n))
        d = rng.standard_normal(n)
        d /= np.linalg.norm(d)
        h = h_scale * max(1.0, np.linalg.norm(x) / np.sqrt(n))
        lam = rng.standard_normal(m)
        sigma = 0.7
        jr, jc = o.jacobianstructure()
        hr, hc = o.hessianstructure()

        def grad_lagrangian(z):
            gL = sigma * np.array(o.gradient(z), dtype=np.float64)
            if m:
                jv = np.array(o.jacobian(z), dtype=np.float64)
                gL = gL + np.bincount(jc, weights=jv * lam[jr], minlength=n)
            return gL

        f0 = float(o.objective(x))
        grad = np.array(o.gradient(x), dtype=np.float64)
        g0 = np.array(o.constraints(x), dtype=np.float64) if m else np.zeros(0)
        jac = np.array(o.jacobian(x), dtype=np.float64)
        hess = np.array(o.hessian(x, lam, sigma), dtype=np.float64)
        assert np.isfinite(f0) and np.isfinite(grad).all() and np.isfinite(jac).all() and np.isfinite(hess).all()
        xp, xm = x + h * d, x - h * d

        def close(a, b, what):
            scale = max(np.linalg.norm(a), np.linalg.norm(b), 1e-300)
            err = np.linalg.norm(a - b) / scale
            assert err < rtol, "%s: relative error %.3e" % (what, err)

        # the objective is a sum of up to millions of terms: use a larger step so that the difference
        # stays above the fp64 cancellation noise (|f| * 1e-16 / h)
        hf = 100 * h
        fd = (float(o.objective(x + hf * d)) - float(o.objective(x - hf * d))) / (2 * hf)
        assert abs(fd - grad @ d) <= 10 * rtol * max(abs(fd), abs(grad @ d), abs(f0) * 1e-16 / hf * 1e3, 1e-300), \
            "df: %r vs %r" % (fd, grad @ d)
        if m:
            gp, gm = np.array(o.constraints(xp), dtype=np.float64), np.array(o.constraints(xm), dtype=np.float64)
            close((gp - gm) / (2 * h), _spmv(jr, jc, jac, d, m), "J d")
        Hd = _spmv(hr, hc, hess, d, n)
        off = hr != hc
        Hd += np.bincount(hc[off], weights=hess[off] * d[hr[off]], minlength=n)     # the upper triangle
        close((grad_lagrangian(xp) - grad_lagrangian(xm)) / (2 * h), Hd, "H d")
        # linearity in the multipliers, sigma = 0 path (Knitro's EVALH_NO_F)
        lam2 = rng.standard_normal(m)
        h1 = hess.copy()
        h2 = np.array(o.hessian(x, lam2, 0.0), dtype=np.float64).copy()
        h12 = np.array(o.hessian(x, lam + lam2, sigma), dtype=np.float64)
        np.testing.assert_allclose(h12, h1 + h2, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(h12).max()))
        # fused evaluation equals the five callbacks
        res = o.eval_all(x, lam, sigma)
        np.testing.assert_allclose(np.array(res["hess"]), hess, rtol=1e-12, atol=0)
        np.testing.assert_allclose(np.array(res["jac"]), jac, rtol=1e-12, atol=0)
        assert float(res["f"]) == f0
    finally:
        o.close()


def test_c2_full_size_eigen_qcqp():
    from dnlp_b200 import workloads as W
    _check_derivatives(W.eigen_qcqp(8192))


def test_c3_logistic_regression():
    from dnlp_b200 import workloads as W
    m, n = (2_000_000, 4096) if FULL else (400_000, 1024)       # full: 34 M Jacobian entries
    At, x0 = W.logistic_data(m, n, 16)
    _check_derivatives(W.logistic_regression(At, x0), rtol=2e-5)


def test_c5_microbench():
    from dnlp_b200 import workloads as W
    N = 10_000_000 if FULL else 800_000                          # full: 50 M nnz, 10 M nodes
    A, x0 = W.microbench_data(N, N // 2, 10)
    _check_derivatives(W.microbench(A, x0), rtol=2e-5)


def test_c4_full_batch_of_4096_starts():
    """BatchedOracles at B = 4096, n = 512, k = 8 (4.3 GB of Hessian values): sampled starts against
    the CPU oracle evaluated start by start, plus a whole-batch invariant that touches every start -
    each start's Hessian is linear in its own (sigma, lambda)."""
    from dnlp_b200 import workloads as W
    from dnlp_b200.multistart import BatchedOracles
    from oracle.dnlp_oracle import RefOracles
    B = 4096 if FULL else 512
    P, q, rng = W.qcqp_data(512, 8)
    prob = W.qcqp(P, q)
    X = rng.uniform(-1, 1, (B, 512))
    lrng = np.random.default_rng(17)
    LAM = lrng.standard_normal((B, 8))
    SIG = lrng.uniform(0.2, 1.5, B)
    o = BatchedOracles(prob, B)
    r = RefOracles(prob)
    jr, jc = r.jacobianstructure()
    hr, hc = r.hessianstructure()
    try:
        np.testing.assert_array_equal(o.jacobianstructure()[0], jr)
        np.testing.assert_array_equal(o.jacobianstructure()[1], jc)
        np.testing.assert_array_equal(o.hessianstructure()[0], hr)
        np.testing.assert_array_equal(o.hessianstructure()[1], hc)
        res = o.eval(X, LAM, SIG)
        for b in sorted({0, 1, 31, 32, 255, 256, 1023, B // 2, B - 2, B - 1} & set(range(B))):
            assert_close(res["f"][b], r.objective(X[b]), "f[%d]" % b)
            assert_close(res["grad"][b], r.gradient(X[b]), "grad[%d]" % b)
            assert_close(res["g"][b], r.constraints(X[b]), "g[%d]" % b)
            assert_close(res["jac"][b], np.asarray(r.jacobian(X[b])).ravel(), "jac[%d]" % b)
            assert_close(res["hess"][b], np.asarray(r.hessian(X[b], LAM[b], SIG[b])).ravel(), "hess[%d]" % b)
        # the QCQP Hessian does not depend on x: H(2 lam, 2 sigma) = 2 H(lam, sigma) for EVERY start
        h1 = res["hess"]
        del res
        h2 = o.eval(X[::-1].copy(), 2.0 * LAM, 2.0 * SIG, want=("hess",))["hess"]
        np.testing.assert_allclose(h2, 2.0 * h1, rtol=1e-12, atol=0)
    finally:
        o.close()
