"""The problems of the reference's OWN NLP test-suite (cvxpy/tests/NLP_tests/test_*.py), harvested by running those
test functions unmodified and recording the problem at their first ``prob.solve(nlp=True, ...)``
(tests/golden/make_golden_refproblems.py; 90 reference tests, 79 distinct problems).  For each: the reference's reduction
chain + ``Oracles`` gave the structures and the five outputs at several points; the oracle port and the DAG compiler
(tape semantics through the test-only NumPy interpreter) must reproduce them - structures and their order bit for
bit, values rel 1e-10.  CPU tier; the GPU tier runs the hand-written set of tests/golden/*.npz through the C-ABI."""
import numpy as np
import pytest

from dnlp_b200.compiler import compile_problem
from golden_util import REFPROBLEMS_DIR, Golden, assert_close, refproblem_golden_names
from oracle.dnlp_oracle import RefOracles
from tape_interp import TapeInterp


def test_every_harvested_problem_is_there():
    assert len(refproblem_golden_names()) == 79


@pytest.mark.parametrize("name", refproblem_golden_names())
def test_oracle_port_on_the_reference_suites_problems(name):
    g = Golden(name, REFPROBLEMS_DIR)
    o = RefOracles(g.problem)
    jr, jc = o.jacobianstructure()
    hr, hc = o.hessianstructure()
    assert jr.dtype == np.int32 and hr.dtype == np.int32
    np.testing.assert_array_equal(jr, g.jac_rows)
    np.testing.assert_array_equal(jc, g.jac_cols)
    np.testing.assert_array_equal(hr, g.hess_rows)
    np.testing.assert_array_equal(hc, g.hess_cols)
    with np.errstate(all="ignore"):
        for i, p in enumerate(g.points):
            assert_close(o.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            if g.problem.m:
                assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i, atol=1e-9)    # rows that cancel to ~0
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)


@pytest.mark.parametrize("name", refproblem_golden_names())
def test_compiler_on_the_reference_suites_problems(name):
    g = Golden(name, REFPROBLEMS_DIR)
    tape = compile_problem(g.problem)
    assert tape.jac_rows.dtype == np.int32 and tape.hess_rows.dtype == np.int32
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.jac_cols, g.jac_cols)
    np.testing.assert_array_equal(tape.hess_rows, g.hess_rows)
    np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    it = TapeInterp(tape)
    for i, p in enumerate(g.points):
        assert_close(it.eval("f", p["x"]), p["f"], "f[%d]" % i)
        assert_close(it.eval("grad", p["x"]), p["grad"], "grad[%d]" % i)
        assert_close(it.eval("g", p["x"]), p["g"], "g[%d]" % i, atol=1e-9)            # rows that cancel to ~0
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac[%d]" % i)
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
