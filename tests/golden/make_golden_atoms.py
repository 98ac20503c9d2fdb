"""Atom-rule golden vectors from the LIVE reference (run in the build container only).

    python tests/golden/make_golden_atoms.py        # writes tests/golden/atoms/*.npz

The reference's tier-1 tests (cvxpy/tests/NLP_tests/jacobian_tests/, hess_tests/) call
``expr.jacobian()`` / ``expr.hess_vec(vec)`` on raw expressions.  Here the same kinds of expressions
are placed, WITHOUT running Dnlp2Smooth, into ``Problem(Minimize(obj), [expr == 0, ...])`` and sent
through the reference's own ``Bounds`` + ``Oracles`` (nlp_solver.py:81-427), so that every rule is
exercised through the seven callbacks with its triplet ORDER recorded (the reference's unit tests
scatter into dense matrices and therefore do not pin order).  Cases the reference rejects are
stored with the exception type; the compiler must reject them with the same type.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import cp, eval_points  # noqa: E402  (loads the reference)

from cvxpy.reductions.solvers.nlp_solvers.nlp_solver import Bounds, Oracles  # noqa: E402

from dnlp_b200 import ir  # noqa: E402
from dnlp_b200.frontend_cvxpy import problem_to_ir  # noqa: E402

OUT = os.path.join(HERE, "atoms")
CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def V(shape, lo=0.3, hi=1.7, seed=0, name=None):
    v = cp.Variable(shape, name=name)
    v.value = np.random.default_rng(seed).uniform(lo, hi, shape if shape != () else None)
    return v


# ---- elementwise atoms on vectors and matrices ----------------------------------------------------
def _elementwise_case(atom, lo=0.3, hi=0.9):
    def build():
        x, X = V(4, lo, hi, 1), V((2, 3), lo, hi, 2)
        return cp.sum(atom(x)), [atom(x) == 0, atom(X) == 0]
    return build


for _name, _atom, _lo, _hi in [
        ("exp", cp.exp, -1, 1), ("log", cp.log, 0.3, 2), ("entr", cp.entr, 0.3, 2), ("logistic", cp.logistic, -2, 2),
        ("sin", cp.sin, -1, 1), ("cos", cp.cos, -1, 1), ("tan", cp.tan, -1, 1), ("sinh", cp.sinh, -1, 1),
        ("tanh", cp.tanh, -1, 1), ("asinh", cp.asinh, -1, 1), ("atanh", cp.atanh, -0.8, 0.8),
        ("xexp", cp.xexp, 0.1, 1), ("square", cp.square, -1, 1), ("cube", lambda v: cp.power(v, 3), -1, 1),
        ("sqrt", lambda v: cp.power(v, 0.5), 0.3, 2), ("pow_2p5", lambda v: cp.power(v, 2.5), 0.3, 2),
        ("pow_third", lambda v: cp.power(v, 1.0 / 3), 0.3, 2)]:
    CASES["elem_" + _name] = _elementwise_case(_atom, _lo, _hi)


# ---- indexing ----------------------------------------------------------------------------------------
@case
def index_int_and_slices():
    x = V(5, 0.3, 2, 3)
    return cp.sum(cp.log(x)[1:4]), [cp.log(x)[1] == 0, cp.log(x)[1:3] == 0, cp.exp(x)[::2] == 0, cp.sin(x)[::-1] == 0,
                                    cp.log(x)[-2:] == 0]


@case
def index_matrix_affine_arg():
    X = V((3, 4), 0.3, 2, 4)
    return 0, [X[0, :] == 0, X[1:, ::2] == 0, X[:, 1] == 0, (2 * X)[2, 3] == 0, X[::-1, 0] == 0, X.T[1:3, :] == 0]


@case
def index_matrix_nonlinear_arg_jac_only_rows():
    X = V((3, 2), 0.3, 2, 5)
    return 0, [cp.log(X)[0, :] == 0]          # hess_vec scatters on a flat vector: fails for 2-D keys (quirk Q6)


@case
def special_index_bool_and_lists():
    x = V(4, 0.3, 2, 6)
    return cp.sum(cp.log(x)[[True, False, True, False]]), [cp.log(x)[[True, False, True, False]] == 0,
                                                            cp.log(x)[[2, 0, 3]] == 0, cp.exp(x)[[0, 0]] == 0]


@case
def special_index_matrix():
    X = V((3, 2), 0.3, 2, 7)
    return 0, [cp.log(X)[[True, False, True], :] == 0, cp.log(X)[[2, 1], [0, 1]] == 0, X[[0, 1], [1, 0]] == 0]


@case
def index_newaxis():
    x = V(3, 0.3, 2, 8)
    return 0, [cp.log(x)[:, None] == 0]


# ---- transpose / reshape / sum / broadcast / promote ----------------------------------------------------
@case
def transpose_cases():
    X = V((2, 3), 0.3, 2, 9)
    A = np.arange(6.0).reshape(3, 2) + 1
    return cp.sum(cp.log(X).T), [cp.log(X).T == 0, cp.sum(cp.log(X), axis=1).T == 0, (A @ cp.log(X)).T == 0, X.T == 0]


@case
def reshape_F():
    x = V(6, 0.3, 2, 10)
    return 0, [cp.reshape(cp.exp(x), (2, 3), order='F') == 0, cp.reshape(x, (3, 2), order='F') == 0]


@case
def reshape_C_rejected():
    x = V(6, 0.3, 2, 10)
    return 0, [cp.reshape(cp.exp(x), (2, 3), order='C') == 0]


@case
def sum_axes():
    X = V((3, 4), 0.3, 2, 11)
    return cp.sum(cp.exp(X)), [cp.sum(cp.exp(X), axis=0) == 0, cp.sum(cp.log(X), axis=1) == 0,
                               cp.sum(cp.sin(X), axis=0, keepdims=True) == 0, cp.sum(X, axis=1, keepdims=True) == 0,
                               cp.sum(cp.exp(X)) == 0]


@case
def broadcast_row_col_scalar():
    r, c, s = V((1, 3), 0.3, 2, 12), V((2, 1), 0.3, 2, 13), V((1, 1), 0.3, 2, 14)
    M = np.arange(6.0).reshape(2, 3) + 1
    return 0, [cp.exp(r) + M == 0, cp.log(c) + M == 0, cp.sin(s) + M == 0, r + c == 0]


@case
def promote_scalar():
    s, x = V((), 0.3, 2, 15), V(3, 0.3, 2, 16)
    return cp.exp(s), [cp.exp(s) + x == 0, s + cp.log(x) == 0, cp.power(s, 2) * np.ones(3) + x == 0]


# ---- products ---------------------------------------------------------------------------------------------
@case
def multiply_const_and_vars():
    x, y = V(4, 0.3, 2, 17), V(4, 0.3, 2, 18)
    c = np.array([1.0, -2.0, 0.0, 3.0])
    return cp.sum(cp.multiply(x, y)), [cp.multiply(c, cp.exp(x)) == 0, cp.multiply(cp.log(x), c) == 0,
                                        cp.multiply(x, y) == 0, cp.multiply(y, x) == 0]


@case
def multiply_matrix_vars():
    X, Y = V((2, 3), 0.3, 2, 19), V((2, 3), 0.3, 2, 20)
    return 0, [cp.multiply(X, Y) == 0]


@case
def multiply_promoted_scalar_var():
    s, x = V((), 0.3, 2, 21), V(4, 0.3, 2, 22)
    return 0, [cp.multiply(s, x) == 0, cp.multiply(x, s) == 0]


@case
def multiply_same_variable_rejected():
    x = V(3, 0.3, 2, 23)
    return 0, [cp.multiply(x, x) == 0]


@case
def matmul_const_left_right():
    Y, X = V((3, 2), 0.3, 2, 24), V((2, 3), 0.3, 2, 25)
    A = np.array([[1.0, 2.0, 3.0], [4.0, 0.0, 6.0]])
    B = np.array([[1.0, 2.0], [0.0, 4.0], [5.0, 6.0]])
    return 0, [A @ Y == 0, X @ B == 0, A @ cp.log(Y) == 0, cp.exp(X) @ B == 0]


@case
def matmul_vector_operands():
    x = V(3, 0.3, 2, 26)
    A = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    return 0, [A @ x == 0, A @ cp.log(x) == 0]


@case
def matmul_both_variables():
    X, Y = V((2, 3), 0.3, 2, 27), V((3, 2), 0.3, 2, 28)
    return cp.sum(X @ Y), [X @ Y == 0, cp.sum(X @ Y) == 0]


@case
def matmul_n1():
    Y = V((1, 3), 0.3, 2, 29)
    return 0, [np.array([[1.0], [2.0], [3.0]]) @ Y == 0]


@case
def matmul_var_times_atom():
    X, Y = V((2, 3), 0.3, 2, 30), V((3, 2), 0.3, 2, 31)
    return 0, [X @ cp.log(Y) == 0]            # Jacobian only: the Hessian rule rejects var @ atom


@case
def matmul_atom_times_var():
    X, Y = V((2, 3), 0.3, 2, 32), V((3, 2), 0.3, 2, 33)
    return 0, [cp.exp(X) @ Y == 0]


@case
def matmul_atom_times_atom():
    X, Y, Z = V((2, 3), 0.3, 2, 34), V((3, 2), 0.3, 2, 35), V((3, 2), 0.3, 2, 36)
    return 0, [cp.exp(X) @ cp.sin(Y) == 0, cp.exp(X) @ (cp.sin(Y) + cp.cos(Z)) == 0]


@case
def matmul_shared_variable_rejected():
    X = V((2, 2), 0.3, 2, 37)
    return 0, [X @ X == 0]


@case
def matmul_sandwich():
    x = V((2, 2), 0.3, 2, 38)
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    return 0, [A @ cp.log(x) @ A == 0]


# ---- multi-argument smooth atoms ---------------------------------------------------------------------------
@case
def rel_entr_three_shapes():
    x, y, s, t = V(3, 0.3, 2, 39), V(3, 0.3, 2, 40), V((), 0.3, 2, 41), V((), 0.3, 2, 42)
    return cp.sum(cp.rel_entr(x, y)), [cp.rel_entr(x, y) == 0, cp.rel_entr(x, s) == 0, cp.rel_entr(t, y) == 0]


@case
def quad_over_lin_and_quad_form():
    x, y = V(3, 0.3, 2, 43), V((), 0.5, 2, 44)
    Q = np.array([[2.0, 0.5, 0.0], [0.5, 1.0, 0.3], [0.0, 0.3, 3.0]])
    return cp.quad_form(x, Q), [cp.quad_over_lin(x, y) == 0, cp.quad_form(x, Q) == 1]


@case
def neg_add_duplicates():
    x = V(3, 0.3, 2, 45)
    return -cp.sum(cp.exp(x)) + cp.sum(cp.exp(x)), [cp.exp(x) + cp.exp(x) - cp.log(x) + x - 2 * x == 0,
                                                  -(-cp.sin(x)) == 0]


@case
def power_zero_and_one():
    x = V(3, 0.3, 2, 46)
    return 0, [cp.power(x, 1) + cp.exp(x) == 0, cp.power(x, 0) + cp.exp(x) == 0]


def make_one(name, fn):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        obj, cons = fn()
        prob = cp.Problem(cp.Minimize(obj), cons)
        bounds = Bounds(prob)
        oracles = Oracles(bounds.new_problem, bounds.x0, len(bounds.cl))
        out = {}
        stage = "jacobianstructure"
        try:
            pir = problem_to_ir(bounds.new_problem, bounds.cl, bounds.cu, bounds.lb, bounds.ub, bounds.x0)
            js, arrays = ir.dump_problem(pir)
            out["ir_json"] = np.array(js)
            for k, v in arrays.items():
                out["ir_" + k] = v
        except Exception as e:          # the frontend itself may reject (unsupported atom)
            out["ir_json"] = np.array("")
            out["frontend_error"] = np.array(type(e).__name__)
        out["jac_error"] = np.array("")
        out["hess_error"] = np.array("")
        try:
            jr, jc = oracles.jacobianstructure()
            out["jac_rows"], out["jac_cols"] = np.asarray(jr, np.int32), np.asarray(jc, np.int32)
        except Exception as e:
            out["jac_error"] = np.array(type(e).__name__)
        stage = "hessianstructure"
        try:
            hr, hc = oracles.hessianstructure()
            out["hess_rows"], out["hess_cols"] = np.asarray(hr, np.int32), np.asarray(hc, np.int32)
        except Exception as e:
            out["hess_error"] = np.array(type(e).__name__)
        del stage
        data = {"x0": bounds.x0, "lb": bounds.lb, "ub": bounds.ub}
        rng = np.random.default_rng(sum(ord(c) * (i + 1) for i, c in enumerate(name)))
        pts = eval_points(data, rng, k=2)
        m = len(bounds.cl)
        out["npoints"] = np.array(len(pts))
        for i, x in enumerate(pts):
            lam = rng.standard_normal(m)
            sigma = float(rng.uniform(0.5, 1.5))
            out["x_%d" % i], out["lam_%d" % i], out["sigma_%d" % i] = x, lam, np.array(sigma)
            with np.errstate(all="ignore"):
                out["f_%d" % i] = np.asarray(oracles.objective(x), np.float64).reshape(-1)[:1]
                out["g_%d" % i] = np.asarray(oracles.constraints(x), np.float64).reshape(-1)
                if not str(out["jac_error"]):
                    out["grad_%d" % i] = np.array(oracles.gradient(x), np.float64).reshape(-1).copy()
                    out["jac_%d" % i] = np.array(oracles.jacobian(x), np.float64).reshape(-1)
                if not str(out["hess_error"]):
                    out["hess_%d" % i] = np.array(oracles.hessian(x, lam, sigma), np.float64).reshape(-1)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return str(out["jac_error"]) or "ok", str(out["hess_error"]) or "ok", str(out.get("frontend_error", ""))


if __name__ == "__main__":
    for name, fn in CASES.items():
        try:
            print("%-44s jac=%-22s hess=%-22s %s" % ((name,) + make_one(name, fn)))
        except Exception as e:
            print("%-44s GENERATOR ERROR %s: %s" % (name, type(e).__name__, e))
