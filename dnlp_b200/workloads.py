"""The five BASELINE.json configurations as smooth-problem IR, built without the reference.

Each builder emits exactly the tree that the reference's own reduction chain
(FlipObjective -> CvxAttr2Constr -> Dnlp2Smooth -> Bounds;
cvxpy/problems/problem.py:1219-1243, reductions/dnlp2smooth/, nlp_solver.py:81-178) produces for
the corresponding CVXPY problem; ``tests/test_workloads.py`` pins that by rebuilding the small
instances stored under ``tests/golden/`` (which came from the live reference) and comparing
structures and values.  Data generators follow SURVEY.md section 8(d).
"""
import numpy as np
import scipy.sparse as sp

from . import ir
from .ir import Node


def _c(v):
    return ir.Constant(v)


def _sum(x):
    return Node("sum", [x], (), axis=None, keepdims=False)


def _add(args, shape):
    return Node("add", args, shape)


def _neg(x):
    return Node("neg", [x], x.shape)


def _pow(x, p):
    return Node("power", [x], x.shape, p=float(p), p_rational=ir.rational_power(p))


# ---------------------------------------------------------------------------------------------
# C1 / C2: maximize quad_form(x, A) s.t. sum_squares(x) == 1     (README.md:26-54 of the reference)
# ---------------------------------------------------------------------------------------------
def eigen_qcqp_data(n, seed=0):
    np.random.seed(seed)
    A = np.random.randn(n, n)
    return A.T @ A


def eigen_qcqp(n, A=None):
    """FlipObjective turns Maximize into minimize -(x'Ax); sum_squares canonicalises through
    quad_over_lin_canon (constant denominator) to 1/1.0 * Sum(power(x, 2)); lower_equality
    subtracts the right-hand side."""
    A = eigen_qcqp_data(n) if A is None else A
    x = ir.Variable(n)
    obj = _neg(Node("quad_form", [x, _c(A)], ()))
    con = _add([Node("multiply", [_c(1.0), _sum(_pow(x, 2))], ()), _c(-1.0)], ())
    return ir.ProblemIR(obj, [con], [x], cl=[0.0], cu=[0.0],
                        lb=np.full(n, -np.inf), ub=np.full(n, np.inf), x0=np.ones(n))


# ---------------------------------------------------------------------------------------------
# C3: nonconvex sparse logistic-type regression
# ---------------------------------------------------------------------------------------------
def distinct_columns(rng, m, n, k):
    """(m, k) int array, every row k distinct uniform columns (vectorised rejection)."""
    cols = rng.integers(0, n, size=(m, k))
    while True:
        cols.sort(axis=1)
        dup = np.zeros(cols.shape, dtype=bool)
        dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
        nd = int(dup.sum())
        if nd == 0:
            return cols
        cols[dup] = rng.integers(0, n, size=nd)


def logistic_data(m, n, k=16, seed=0):
    rng = np.random.default_rng(seed)
    cols = distinct_columns(rng, m, n, k).reshape(-1)
    vals = rng.standard_normal(m * k)
    y = rng.choice([-1.0, 1.0], m)
    vals *= np.repeat(-y, k)                       # A~ = diag(-y) A
    indptr = np.arange(0, m * k + 1, k, dtype=np.int64)
    At = sp.csr_array((vals, cols, indptr), shape=(m, n))
    x0 = 0.1 * rng.standard_normal(n)
    return At, x0


def logistic_regression(At, x_init):
    """minimize sum(logistic(A~ x)) + 0.1 sum(log(1 + x^2)) + 0.01 sum(exp(-x)).

    Dnlp2Smooth lifts each nonlinear atom whose argument is not a bare variable
    (logistic_canon.py, log_canon.py, exp_canon.py): t1 == A~x, t2 == 1 + x^2 (t2 >= 0), t3 == -x.
    Variable order is first appearance: [t1, t2, t3, x]  (SURVEY quirk Q1)."""
    m, n = At.shape
    t1, t2, t3 = ir.Variable(m), ir.Variable(n), ir.Variable(n)
    x = ir.Variable(n)
    obj = _add([
        _sum(Node("logistic", [t1], (m,))),
        Node("multiply", [_c(0.1), _sum(Node("log", [t2], (n,)))], ()),
        Node("multiply", [_c(0.01), _sum(Node("exp", [t3], (n,)))], ()),
    ], ())
    c1 = _add([t1, _neg(Node("matmul", [_c(sp.csr_array(At)), x], (m,)))], (m,))
    c2 = _add([t2, _neg(_add([_c(np.ones(n)), _pow(x, 2)], (n,)))], (n,))
    c3 = _add([t3, _neg(_neg(x))], (n,))
    N = m + 3 * n
    lb = np.full(N, -np.inf)
    lb[m:m + n] = 0.0
    x_init = np.asarray(x_init, dtype=np.float64)
    x0 = np.concatenate([At @ x_init, 1 + x_init ** 2, -x_init, x_init])
    return ir.ProblemIR(obj, [c1, c2, c3], [t1, t2, t3, x], cl=np.zeros(m + 2 * n), cu=np.zeros(m + 2 * n),
                        lb=lb, ub=np.full(N, np.inf), x0=x0)


# ---------------------------------------------------------------------------------------------
# C4: nonconvex QCQP (one start; the multi-start batch shards start points across GPUs)
# ---------------------------------------------------------------------------------------------
def qcqp_data(n, k, seed=0):
    rng = np.random.default_rng(seed)
    P = []
    for _ in range(k + 1):
        G = rng.standard_normal((n, n))
        P.append((G + G.T) / 2)
    q = rng.standard_normal((k + 1, n))
    return P, q, rng


def qcqp(P, q, x0=None):
    """minimize x'P0x + q0'x  s.t.  x'Pix + qi'x <= 1, -1 <= x <= 1.
    lower_ineq_to_nonneg gives 1 - (x'Pix + qi'x) >= 0 (reductions/utilities.py:36-39)."""
    n = P[0].shape[0]
    k = len(P) - 1
    x = ir.Variable(n)

    def quad(i):
        return _add([Node("quad_form", [x, _c(P[i])], ()), Node("matmul", [_c(q[i]), x], ())], ())
    cons = [_add([_c(1.0), _neg(quad(i))], ()) for i in range(1, k + 1)]
    return ir.ProblemIR(quad(0), cons, [x], cl=np.zeros(k), cu=np.full(k, np.inf),
                        lb=-np.ones(n), ub=np.ones(n), x0=np.zeros(n) if x0 is None else x0)


# ---------------------------------------------------------------------------------------------
# C5: synthetic oracle microbenchmark: N elementwise nodes + nnz-heavy CSR constraint Jacobian
# ---------------------------------------------------------------------------------------------
C5_OPS = ("exp", "logistic", "sin", "cos", "tanh", "sinh", ("power", 2), ("power", 3))


def microbench_data(N, m, k=10, seed=0):
    rng = np.random.default_rng(seed)
    cols = distinct_columns(rng, m, N, k).reshape(-1)
    vals = rng.standard_normal(m * k)
    indptr = np.arange(0, m * k + 1, k, dtype=np.int64)
    A = sp.csr_array((vals, cols, indptr), shape=(m, N))
    x0 = rng.uniform(0.5, 1.5, N)
    return A, x0


def microbench(A, x0, ops=C5_OPS):
    """minimize sum_s 1'phi_s(x_s)  s.t.  sum_s A_s phi_s(x_s) == 0, one Variable per segment.
    These atoms' canonicalisers leave a bare-Variable argument alone, so the smooth problem has
    exactly N variables, nnzJ = nnz(A), nnzH = N."""
    m, N = A.shape
    S = len(ops)
    seg = N // S
    assert seg * S == N
    A = sp.csc_array(A)
    xs = [ir.Variable(seg) for _ in ops]

    def phi(op, v):
        if isinstance(op, tuple):
            return _pow(v, op[1])
        return Node(op, [v], v.shape)
    obj = _add([_sum(phi(op, v)) for op, v in zip(ops, xs)], ())
    terms = [Node("matmul", [_c(sp.csr_array(A[:, s * seg:(s + 1) * seg])), phi(op, v)], (m,))
             for s, (op, v) in enumerate(zip(ops, xs))]
    con = _add(terms + [_c(-np.zeros(m))], (m,))
    return ir.ProblemIR(obj, [con], xs, cl=np.zeros(m), cu=np.zeros(m),
                        lb=np.full(N, -np.inf), ub=np.full(N, np.inf), x0=x0)
