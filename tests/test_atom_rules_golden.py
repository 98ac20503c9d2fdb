"""Per-rule parity against the live reference's Oracles on raw (un-canonicalised) expressions:
CPU oracle and DAG compiler, including the cases the reference rejects (same exception type)."""
import builtins

import numpy as np
import pytest

from dnlp_b200.compiler import compile_problem
from golden_util import REFTESTS_DIR, AtomGolden, assert_close, atom_golden_names, reftest_golden_names
from oracle.dnlp_oracle import RefOracles
from tape_interp import TapeInterp


def _exc(name):
    return getattr(builtins, name)


@pytest.mark.parametrize("name", atom_golden_names())
def test_oracle_rules(name):
    _check_oracle_rules(AtomGolden(name))


@pytest.mark.parametrize("name", reftest_golden_names())
def test_oracle_rules_on_the_reference_suites_own_expressions(name):
    """Every expression the reference's jacobian_tests / hess_tests differentiate (136 tests, 85 distinct
    expressions the reference's Oracles can evaluate): tests/golden/make_golden_reftests.py."""
    _check_oracle_rules(AtomGolden(name, REFTESTS_DIR))


def _check_oracle_rules(g):
    o = RefOracles(g.problem)
    if g.jac_error:
        with pytest.raises(_exc(g.jac_error)):
            o.jacobianstructure()
    else:
        jr, jc = o.jacobianstructure()
        np.testing.assert_array_equal(jr, g.jac_rows)
        np.testing.assert_array_equal(jc, g.jac_cols)
    if g.hess_error:
        with pytest.raises(_exc(g.hess_error)):
            o.hessianstructure()
    else:
        hr, hc = o.hessianstructure()
        np.testing.assert_array_equal(hr, g.hess_rows)
        np.testing.assert_array_equal(hc, g.hess_cols)
    with np.errstate(all="ignore"):
        for p in g.points:
            assert_close(o.objective(p["x"]), p["f"], "f")
            assert_close(o.constraints(p["x"]), p["g"], "g")
            if not g.jac_error:
                assert_close(o.gradient(p["x"]), p["grad"], "grad")
                assert_close(o.jacobian(p["x"]), p["jac"], "jac")
            if not g.hess_error:
                assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")


@pytest.mark.parametrize("name", atom_golden_names())
def test_compiler_rules(name):
    _check_compiler_rules(AtomGolden(name))


@pytest.mark.parametrize("name", reftest_golden_names())
def test_compiler_rules_on_the_reference_suites_own_expressions(name):
    _check_compiler_rules(AtomGolden(name, REFTESTS_DIR))


def test_reference_suite_fixtures_are_all_there():
    assert len(reftest_golden_names()) == 85


def _check_compiler_rules(g):
    if g.jac_error:
        with pytest.raises(_exc(g.jac_error)):
            compile_problem(g.problem, with_hessian=False)
        return
    if g.hess_error:
        with pytest.raises(_exc(g.hess_error)):
            compile_problem(g.problem)
        tape = compile_problem(g.problem, with_hessian=False)
    else:
        tape = compile_problem(g.problem)
        np.testing.assert_array_equal(tape.hess_rows, g.hess_rows)
        np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.jac_cols, g.jac_cols)
    it = TapeInterp(tape)
    for p in g.points:
        assert_close(it.eval("f", p["x"]), p["f"], "f")
        assert_close(it.eval("g", p["x"]), p["g"], "g")
        assert_close(it.eval("grad", p["x"]), p["grad"], "grad")
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac")
        if not g.hess_error:
            assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")
