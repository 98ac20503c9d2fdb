"""Drive the UNMODIFIED reference (oracle/_ref, built by oracle/make_ref.sh) on the bench workloads.

TEST INFRASTRUCTURE / CPU BASELINE ONLY: imported by ``bench.py``'s ``--impl reference`` arm and its
``cpu_baseline`` leg, and by tests.  Nothing under ``dnlp_b200/`` imports this module.

Each workload is written as the CVXPY problem a user of the reference would write (SURVEY.md section
8d), on the SAME seeded data generators as ``dnlp_b200.workloads``; the reference's own reduction chain
(cvxpy/problems/problem.py:1219-1243) and its ``Oracles`` (nlp_solver.py:181-427) do the rest.
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "cvxpy", "version.py"))


def load_reference():
    """import cvxpy from oracle/_ref (raises if make_ref.sh has not been run)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run oracle/make_ref.sh where /root/reference exists")
    if "cvxpy" in sys.modules:
        mod = sys.modules["cvxpy"]
        if not os.path.abspath(mod.__file__).startswith(REF_DIR):
            raise RuntimeError("another cvxpy is already imported from %s" % mod.__file__)
        return mod
    sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    import cvxpy
    return cvxpy


def reference_data(prob):
    """The reference's nlp chain up to the Oracles object (problem.py:1220-1243)."""
    cp = load_reference()
    from cvxpy.reductions.cvx_attr2constr import CvxAttr2Constr
    from cvxpy.reductions.dnlp2smooth.dnlp2smooth import Dnlp2Smooth
    from cvxpy.reductions.flip_objective import FlipObjective
    from cvxpy.reductions.solvers.nlp_solvers.ipopt_nlpif import IPOPT
    from cvxpy.reductions.solvers.solving_chain import SolvingChain
    assert prob.is_dnlp()
    red = ([FlipObjective()] if type(prob.objective) == cp.Maximize else []) + \
        [CvxAttr2Constr(reduce_bounds=False), Dnlp2Smooth(), IPOPT()]
    data, _ = SolvingChain(reductions=red).apply(problem=prob)
    return data


# ---------------------------------------------------------------------------------------------------
# the bench workloads as CVXPY problems
# ---------------------------------------------------------------------------------------------------
def eigen_qcqp(n):
    cp = load_reference()
    from dnlp_b200 import workloads as W
    A = W.eigen_qcqp_data(n)
    x = cp.Variable(n)
    x.value = np.ones(n)
    return cp.Problem(cp.Maximize(cp.quad_form(x, A, assume_PSD=True)), [cp.sum_squares(x) == 1])


def logistic_regression(m, n, k=16):
    cp = load_reference()
    from dnlp_b200 import workloads as W
    At, x0 = W.logistic_data(m, n, k)
    x = cp.Variable(n)
    x.value = x0
    obj = cp.sum(cp.logistic(sp.csr_matrix(At) @ x)) + 0.1 * cp.sum(cp.log(1 + cp.power(x, 2))) \
        + 0.01 * cp.sum(cp.exp(-x))
    return cp.Problem(cp.Minimize(obj))


def qcqp(n=512, k=8, start=0):
    cp = load_reference()
    from dnlp_b200 import workloads as W
    P, q, rng = W.qcqp_data(n, k)
    X = rng.uniform(-1, 1, (max(start + 1, 1), n))
    x = cp.Variable(n, bounds=[-1, 1])
    x.value = X[start]
    cons = [cp.quad_form(x, P[i], assume_PSD=True) + q[i] @ x <= 1 for i in range(1, k + 1)]
    return cp.Problem(cp.Minimize(cp.quad_form(x, P[0], assume_PSD=True) + q[0] @ x), cons)


def microbench(N, m=None, k=10):
    cp = load_reference()
    from dnlp_b200 import workloads as W
    m = N // 2 if m is None else m
    A, x0 = W.microbench_data(N, m, k)
    ops = [cp.exp, cp.logistic, cp.sin, cp.cos, cp.tanh, cp.sinh,
           lambda v: cp.power(v, 2), lambda v: cp.power(v, 3)]          # = workloads.C5_OPS
    seg = N // len(ops)
    xs = [cp.Variable(seg) for _ in ops]
    for s, v in enumerate(xs):
        v.value = x0[s * seg:(s + 1) * seg]
    Ac = sp.csc_matrix(A)
    g, f = 0, 0
    for s, (op, v) in enumerate(zip(ops, xs)):
        g = g + sp.csr_matrix(Ac[:, s * seg:(s + 1) * seg]) @ op(v)
        f = f + cp.sum(op(v))
    return cp.Problem(cp.Minimize(f), [g == 0])


def build(workload, scale):
    """(cvxpy problem, size description, work units) for a bench workload at ``scale`` of its full size.
    Work units = Jacobian + 2 x Hessian triplets + n + m after the structure passes (the reference's
    per-call cost is linear in them)."""
    if workload == "c1":
        return eigen_qcqp(3), "n=3"
    if workload == "c2":
        n = max(8, int(round(8192 * scale)))
        return eigen_qcqp(n), "n=%d" % n
    if workload in ("c3", "c3s"):
        m = max(64, int(2_000_000 * scale))
        return logistic_regression(m, 4096), "m=%d n=4096" % m
    if workload == "c4":
        return qcqp(512, 8), "n=512 k=8, one start"
    if workload in ("c5", "c5s"):
        N = max(64, int(10_000_000 * scale) // 8 * 8)
        return microbench(N), "N=%d nnz=%d" % (N, N // 2 * 10)
    raise KeyError(workload)


class TimedReference:
    """The reference's Oracles on one problem; ``step()`` = the five callbacks at a fresh point."""

    def __init__(self, prob):
        t0 = time.perf_counter()
        self.data = reference_data(prob)
        self.o = self.data["oracles"]
        self.chain_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.jr, _ = self.o.jacobianstructure()          # structure FIRST, as IPOPT does
        self.hr, _ = self.o.hessianstructure()
        self.structure_s = time.perf_counter() - t0
        self.n, self.m = self.data["x0"].size, len(self.data["cl"])
        self.work = len(self.jr) + 2 * len(self.hr) + self.n + self.m
        self.rng = np.random.default_rng(1000)
        self.x0 = np.asarray(self.data["x0"], np.float64)

    def step(self):
        x = self.x0 * (1.0 + 0.01 * self.rng.standard_normal(self.n))
        lam = self.rng.standard_normal(self.m)
        o = self.o
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            o.objective(x), o.gradient(x), o.constraints(x), o.jacobian(x), o.hessian(x, lam, 1.0)
        return time.perf_counter() - t0
