"""The native IR builders for the BASELINE configs reproduce what the reference's own reduction
chain emits (checked against the golden fixtures, which were produced by the live reference)."""
import numpy as np
import scipy.sparse as sp

from dnlp_b200 import workloads as W
from dnlp_b200.compiler import compile_problem
from golden_util import Golden, assert_close
from tape_interp import TapeInterp


def _check(prob, g):
    tape = compile_problem(prob)
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.jac_cols, g.jac_cols)
    np.testing.assert_array_equal(tape.hess_rows, g.hess_rows)
    np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    it = TapeInterp(tape)
    for p in g.points:
        assert_close(it.eval("f", p["x"]), p["f"], "f")
        assert_close(it.eval("grad", p["x"]), p["grad"], "grad")
        assert_close(it.eval("g", p["x"]), p["g"], "g")
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac")
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")


def test_c1_toy_matches_reference_chain():
    g = Golden("c1_readme_toy")
    prob = W.eigen_qcqp(3)
    np.testing.assert_array_equal(prob.x0, g.problem.x0)
    _check(prob, g)


def test_c2_matches_reference_chain():
    _check(W.eigen_qcqp(24), Golden("c2_eigen_qcqp_small"))


def test_c3_matches_reference_chain():
    rng = np.random.default_rng(0)                      # same draws as tests/golden/make_golden.py
    m, n, k = 200, 16, 4
    cols = np.concatenate([rng.choice(n, k, replace=False) for _ in range(m)])
    A = sp.csr_matrix((rng.standard_normal(m * k), (np.repeat(np.arange(m), k), cols)), shape=(m, n))
    y = rng.choice([-1.0, 1.0], m)
    At = sp.diags(-y) @ A
    x_init = 0.1 * rng.standard_normal(n)
    g = Golden("c3_logistic_small")
    prob = W.logistic_regression(sp.csr_array(At), x_init)
    np.testing.assert_array_equal(prob.lb, g.problem.lb)
    assert_close(prob.x0, g.problem.x0, "x0")
    _check(prob, g)


def test_c4_matches_reference_chain():
    P, q, rng = W.qcqp_data(12, 3)
    _check(W.qcqp(P, q), Golden("c4_qcqp_small"))


def test_c5_matches_reference_chain():
    rng = np.random.default_rng(0)
    seg, m, k = 5, 20, 10
    N = seg * 8
    x0 = np.concatenate([rng.uniform(0.5, 1.5, seg) for _ in range(8)])
    cols = np.concatenate([rng.choice(N, k, replace=False) for _ in range(m)])
    A = sp.csr_matrix((rng.standard_normal(m * k), (np.repeat(np.arange(m), k), cols)), shape=(m, N))
    g = Golden("c5_microbench_small")
    prob = W.microbench(sp.csr_array(A), x0)
    np.testing.assert_array_equal(prob.x0, g.problem.x0)
    _check(prob, g)


def test_distinct_columns_are_distinct():
    cols = W.distinct_columns(np.random.default_rng(3), 500, 40, 16)
    assert cols.shape == (500, 16)
    assert all(len(set(r)) == 16 for r in cols)
