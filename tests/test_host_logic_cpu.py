"""Host logic of ``GpuOracles`` in the CPU tier: the same checks tests/test_gpu_parity.py makes on a B200, with the
tape executed by the NumPy interpreter behind the library's entry-point names (tests/host_logic_device.py, test
infrastructure).  What is under test here is the Python side: constant-entry elision with compact dynamic transfers,
sigma-keyed Hessian entries, reused buffers and the reference's return conventions, parameter re-arming, argument
checks - against the live-reference fixtures."""
import builtins

import numpy as np
import pytest

import host_logic_device
from dnlp_b200.oracles import GpuOracles
from golden_util import (REFPROBLEMS_DIR, REFTESTS_DIR, AtomGolden, Golden, assert_close, atom_golden_names, golden_names,
                         refproblem_golden_names, reftest_golden_names)


@pytest.fixture
def oracles(monkeypatch):
    host_logic_device.install(monkeypatch)
    made = []

    def make(problem, **kw):
        o = GpuOracles(problem, **kw)
        made.append(o)
        return o
    yield make
    for o in made:
        o.close()


def _check_golden(o, g, g_atol=1e-12):
    np.testing.assert_array_equal(o.jacobianstructure()[0], g.jac_rows)
    np.testing.assert_array_equal(o.jacobianstructure()[1], g.jac_cols)
    np.testing.assert_array_equal(o.hessianstructure()[0], g.hess_rows)
    np.testing.assert_array_equal(o.hessianstructure()[1], g.hess_cols)
    for p in g.points + g.points[::-1]:                     # forwards and back: every output array is reused
        assert_close(o.objective(p["x"]), p["f"], "f")
        assert_close(o.gradient(p["x"]), p["grad"], "grad")
        assert_close(o.constraints(p["x"]), p["g"], "g", atol=g_atol)
        assert_close(o.jacobian(p["x"]), p["jac"], "jac")
        assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")
        res = o.eval_all(p["x"], p["lam"], float(p["sigma"]))
        for k in ("f", "grad", "g", "jac", "hess"):
            assert_close(res[k], p[k], "eval_all/" + k, atol=g_atol if k == "g" else 1e-12)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("elide", [False, True])
def test_callbacks_match_the_reference_fixtures(name, elide, oracles, monkeypatch):
    """``elide``: thresholds lowered so that even these small problems take the compact path (only the x / lambda
    dependent entries are fetched, constants are written into the reused output once)."""
    if elide:
        monkeypatch.setattr(GpuOracles, "ELIDE_MIN", 1)
        monkeypatch.setattr(GpuOracles, "ELIDE_MAX_FRACTION", 1.0)
    g = Golden(name)
    o = oracles(g.problem)
    if elide:
        assert set(o._dyn) <= {"jac", "hess", "g", "grad"}
    _check_golden(o, g)


@pytest.mark.parametrize("name", refproblem_golden_names())
def test_callbacks_on_the_reference_suites_own_problems(name, oracles, monkeypatch):
    """The 79 problems harvested from the reference's problem-level tests, compact path forced on."""
    monkeypatch.setattr(GpuOracles, "ELIDE_MIN", 1)
    monkeypatch.setattr(GpuOracles, "ELIDE_MAX_FRACTION", 1.0)
    g = Golden(name, REFPROBLEMS_DIR)
    _check_golden(oracles(g.problem), g, g_atol=1e-9)


@pytest.mark.parametrize("name", atom_golden_names())
def test_raw_rules_through_the_oracle_object(name, oracles):
    _check_raw_rules(AtomGolden(name), oracles)


@pytest.mark.parametrize("name", reftest_golden_names())
def test_reference_suite_expressions_through_the_oracle_object(name, oracles):
    _check_raw_rules(AtomGolden(name, REFTESTS_DIR), oracles)


def _check_raw_rules(g, oracles):
    if g.jac_error:
        with pytest.raises(getattr(builtins, g.jac_error)):
            oracles(g.problem, with_hessian=False)
        return
    o = oracles(g.problem, with_hessian=not g.hess_error)
    np.testing.assert_array_equal(o.jacobianstructure()[0], g.jac_rows)
    np.testing.assert_array_equal(o.jacobianstructure()[1], g.jac_cols)
    for p in g.points:
        assert_close(o.objective(p["x"]), p["f"], "f")
        assert_close(o.constraints(p["x"]), p["g"], "g")
        assert_close(o.gradient(p["x"]), p["grad"], "grad")
        assert_close(o.jacobian(p["x"]), p["jac"], "jac")
        if not g.hess_error:
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")
        else:
            with pytest.raises(RuntimeError):
                o.hessian(p["x"], p["lam"], 1.0)


def test_return_conventions_and_argument_checks(oracles):
    """Reference quirk Q2 and the boundary's error behaviour: the gradient is the same array object every
    call, the objective a NumPy scalar, structures int32; wrong sizes raise ValueError; duals may be longer than m."""
    g = Golden("hs071")
    o = oracles(g.problem)
    p = g.points[0]
    f = o.objective(p["x"])
    assert isinstance(f, np.float64)
    assert o.gradient(p["x"]) is o.gradient(g.points[1]["x"]) is o.grad_obj
    assert o.jacobian(p["x"]) is o.jacobian(p["x"])
    assert all(a.dtype == np.int32 for a in o.jacobianstructure() + o.hessianstructure())
    with pytest.raises(ValueError):
        o.objective(p["x"][:-1])
    with pytest.raises(ValueError):
        o.hessian(p["x"], p["lam"][:-1], 1.0)
    longer = np.concatenate([p["lam"], np.full(o.n, 123.0)])          # Knitro: constraint AND bound multipliers
    assert_close(o.hessian(list(p["x"]), list(longer), float(p["sigma"])), p["hess"], "hess with m + n duals")
    o.intermediate(0, 7, 0.0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert o.iterations == 7 and o.num_constraints == o.m


def test_sigma_keyed_hessian_entries_are_refetched_only_when_sigma_changes(oracles, monkeypatch):
    """Dense quad_form objective: the 2*sigma*Q layer depends on sigma only.  With elision active those entries travel
    when sigma differs from the previous call and are skipped otherwise (GpuOracles.hessian)."""
    monkeypatch.setattr(GpuOracles, "ELIDE_MIN", 1)
    g = Golden("c2_eigen_qcqp_small")
    o = oracles(g.problem)
    assert o._hess_sigma_class and "hess" in o._dyn
    log = o.dev._L.calls
    p0, p1 = g.points[0], g.points[1]
    for p, sigma in ((p0, 1.0), (p1, 1.0), (p0, 0.25), (p1, 0.25), (p0, 1.0)):
        want = _reference_hessian(g, p, sigma)
        before = len(log)
        assert_close(o.hessian(p["x"], p["lam"], sigma), want, "hess sigma=%g" % sigma)
        assert log[before:] == ["hess"]
    assert o._hess_sigma == 1.0


def _reference_hessian(g, p, sigma):
    from oracle.dnlp_oracle import RefOracles
    r = RefOracles(g.problem)
    r.jacobianstructure(), r.hessianstructure()
    return r.hessian(p["x"], p["lam"], sigma)


def test_parameter_rearming_through_the_oracle_object(oracles):
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_parameters import _problem
    from oracle.dnlp_oracle import RefOracles
    prob, (gamma, b, w), rng = _problem()
    o = oracles(prob)
    xv, lam = prob.x0 * 1.1, rng.standard_normal(prob.m)
    for trial in range(3):
        if trial:
            gamma.attrs["value"] = np.asarray(rng.uniform(0.1, 3.0))
            b.attrs["value"] = rng.standard_normal(30)
            w.attrs["value"] = rng.uniform(0.5, 2, 8)
            o.rearm(prob)                                  # new values into the slots: no recompile
        r = RefOracles(prob.folded())
        r.jacobianstructure(), r.hessianstructure()
        assert_close(o.objective(xv), r.objective(xv), "f")
        assert_close(o.gradient(xv), r.gradient(xv), "grad")
        assert_close(o.constraints(xv), r.constraints(xv), "g")
        assert_close(o.jacobian(xv), r.jacobian(xv), "jac")
        assert_close(o.hessian(xv, lam, 0.6), r.hessian(xv, lam, 0.6), "hess")
    with pytest.raises(ValueError):
        o.set_parameters(np.zeros(3))
