"""Parameter slots (cvxpy/expressions/constants/parameter.py:35): a Parameter in a sum, an
elementwise product or under an affine atom is a value slot of the tape - new values never recompile.
CPU tier: compiler + NumPy tape interpreter against the CPU oracle of the folded problem."""
import numpy as np
import pytest

from dnlp_b200 import ir
from dnlp_b200 import tape as T
from dnlp_b200.compile_cache import fingerprint
from dnlp_b200.compiler import compile_problem
from oracle.dnlp_oracle import RefOracles
from tape_interp import TapeInterp


def _problem(seed=3, n=30, k=8):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((k, n))
    x, t = ir.Variable(n), ir.Variable(k)
    gamma, b, w = ir.Parameter((), 0.5), ir.Parameter(n, rng.standard_normal(n)), ir.Parameter(k, rng.uniform(0.5, 2, k))
    obj = ir.sum(ir.exp(x)) + ir.multiply(gamma, ir.sum(ir.power(x, 2))) + ir.sum(ir.multiply(b, x))
    cons = [ir.multiply(w, ir.logistic(t)) + ir.neg(ir.promote(gamma, (k,))), t + ir.neg(ir.matmul(A, x)),
            ir.sum(ir.multiply(b, ir.power(x, 3))) + ir.neg(gamma)]
    return ir.ProblemIR(obj, cons, [x, t], x0=0.1 * rng.standard_normal(n + k)), (gamma, b, w), rng


def test_parameter_sweep_matches_folded_oracle():
    prob, (gamma, b, w), rng = _problem()
    assert [q.size for q in prob.params] == [1, 30, 8] and prob.n_params == 39
    tape = compile_problem(prob)
    assert tape.n_params == 39 and tape.tmp_slot == prob.n + 1 + prob.m + 39
    assert any(i.dep_mask & T.DEP_PARAM for i in tape.instrs)
    it = TapeInterp(tape)
    xv, lam = prob.x0 * 1.1, rng.standard_normal(prob.m)
    fp = fingerprint(prob)
    for trial in range(4):
        if trial:
            gamma.attrs["value"] = np.asarray(rng.uniform(0.1, 3.0))
            b.attrs["value"] = rng.standard_normal(30)
            w.attrs["value"] = rng.uniform(0.5, 2, 8)
            it.set_params(prob.param_values())
            assert fingerprint(prob) == fp               # values are not part of the tape's identity
        r = RefOracles(prob.folded())
        np.testing.assert_array_equal(tape.jac_rows, r.jacobianstructure()[0])
        np.testing.assert_array_equal(tape.jac_cols, r.jacobianstructure()[1])
        np.testing.assert_array_equal(tape.hess_rows, r.hessianstructure()[0])
        np.testing.assert_array_equal(tape.hess_cols, r.hessianstructure()[1])
        np.testing.assert_allclose(it.eval("f", xv), r.objective(xv), rtol=1e-12)
        np.testing.assert_allclose(it.eval("grad", xv), r.gradient(xv), rtol=1e-12)
        np.testing.assert_allclose(it.eval("g", xv), r.constraints(xv), rtol=1e-12)
        np.testing.assert_allclose(it.eval("jac", xv), np.asarray(r.jacobian(xv)).ravel(), rtol=1e-12)
        np.testing.assert_allclose(it.eval("hess", xv, lam, 0.6), np.asarray(r.hessian(xv, lam, 0.6)).ravel(), rtol=1e-12)


def test_parameter_as_matrix_of_a_product_is_rejected():
    """A Parameter as the constant MATRIX of a product with variables (or quad_form's matrix) would change
    coefficient arrays: not a slot.  The builder says so (the cvxpy frontend freezes such parameters instead)."""
    x = ir.Variable(3)
    P = ir.Parameter((2, 3), np.arange(6.0).reshape(2, 3))
    with pytest.raises(NotImplementedError):
        compile_problem(ir.ProblemIR(ir.sum(ir.exp(x)), [ir.matmul(P, x) + (-1.0)], [x], x0=np.zeros(3)))
    Q = ir.Parameter((3, 3), np.eye(3))
    with pytest.raises(NotImplementedError):
        compile_problem(ir.ProblemIR(ir.quad_form(x, Q), [], [x], x0=np.ones(3)))


def test_fingerprint_distinguishes_parameter_layout_not_values():
    p1, (g1, b1, w1), _ = _problem(seed=3)
    p2, (g2, b2, w2), _ = _problem(seed=3)
    b2.attrs["value"] = b2.attrs["value"] + 1.0
    assert fingerprint(p1) == fingerprint(p2)
    p3, _, _ = _problem(seed=4)                          # other data matrix: another tape
    assert fingerprint(p3) != fingerprint(p1)
