"""DAG compiler vs the live-reference golden vectors (CPU: patterns bit-exact, tape semantics
checked with the test-only NumPy tape interpreter)."""
import numpy as np
import pytest

from dnlp_b200.compiler import compile_problem
from golden_util import Golden, assert_close, golden_names
from tape_interp import TapeInterp


@pytest.mark.parametrize("name", golden_names())
def test_compiled_tape_matches_reference(name):
    g = Golden(name)
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
        assert_close(it.eval("g", p["x"]), p["g"], "g[%d]" % i)
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac[%d]" % i)
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)


@pytest.mark.parametrize("name", golden_names())
def test_optimised_emission_paths(name, monkeypatch):
    """Same parity with the large-problem code paths forced on: two-stage long-row reduction and
    the streaming first-layer (SCALE / one-term POLY) + scatter-accumulate output split."""
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "LAYER_MIN", 2)
    monkeypatch.setattr(Builder, "LONG_ROW", 3)
    monkeypatch.setattr(Builder, "CHUNK", 2)
    g = Golden(name)
    tape = compile_problem(g.problem)
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    it = TapeInterp(tape)
    for i, p in enumerate(g.points):
        assert_close(it.eval("f", p["x"]), p["f"], "f[%d]" % i)
        assert_close(it.eval("grad", p["x"]), p["grad"], "grad[%d]" % i)
        assert_close(it.eval("g", p["x"]), p["g"], "g[%d]" % i)
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac[%d]" % i)
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
