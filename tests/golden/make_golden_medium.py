"""Medium-size golden vectors of the four big BASELINE configs, from the LIVE reference.

    python tests/golden/make_golden_medium.py          # writes tests/golden/medium/*.npz

The small fixtures (make_golden.py) never reach the production kernels; the full BASELINE sizes need
minutes and tens of GB in the reference's Python lists.  These sit in between: the same data
generators the benchmark uses (``dnlp_b200.workloads``), at sizes the reference finishes in seconds,
evaluated by the reference's own chain and ``Oracles`` (nlp_solver.py:181-427):

    c2_medium   eigen-QCQP n = 1024 (dense quad_form, GEMV + SCALE kernels)
    c3_medium   logistic-type regression m = 20 000, n = 1024, 16 nnz/row
    c4_medium   nonconvex QCQP n = 512, k = 8, 4 of the 4096 multi-start points
    c5_medium   microbenchmark N = 100 000 nodes, m = 50 000 rows, 10 nnz/row

Stored per config: the evaluation point(s), every output in full, and SHA-256 digests of the int32
structure arrays (bit-exact check without carrying 2 x nnz integers).  The problem data is NOT
stored: the tests rebuild it from the same seeded generators.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import cp, reference_data  # noqa: E402  (injects the version stub, imports the reference)

from dnlp_b200 import workloads as W  # noqa: E402

OUT = os.path.join(HERE, "medium")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


def evaluate(prob, points):
    """points: list of (x or None, lam_seed).  x None = the chain's own x0 perturbed."""
    data = reference_data(prob)
    o = data["oracles"]
    jr, jc = o.jacobianstructure()
    hr, hc = o.hessianstructure()
    n, m = data["x0"].size, len(data["cl"])
    out = {"n": n, "m": m, "nnz_jac": len(jr), "nnz_hess": len(hr),
           "jac_rows_sha256": digest(jr), "jac_cols_sha256": digest(jc),
           "hess_rows_sha256": digest(hr), "hess_cols_sha256": digest(hc),
           "x0": np.asarray(data["x0"], np.float64), "npoints": len(points)}
    for i, (x, seed, sigma) in enumerate(points):
        rng = np.random.default_rng(seed)
        if x is None:
            x = np.asarray(data["x0"], np.float64) * (1 + 0.01 * rng.standard_normal(n))
        lam = rng.standard_normal(m)
        with np.errstate(all="ignore"):
            out["x_%d" % i] = x
            out["lam_%d" % i] = lam
            out["sigma_%d" % i] = float(sigma)
            out["f_%d" % i] = float(o.objective(x))
            out["grad_%d" % i] = np.array(o.gradient(x), np.float64).ravel()
            out["g_%d" % i] = np.array(o.constraints(x), np.float64).ravel()
            out["jac_%d" % i] = np.array(o.jacobian(x), np.float64).ravel()
            out["hess_%d" % i] = np.array(o.hessian(x, lam, sigma), np.float64).ravel()
    return out


def c2_medium():
    n = 1024
    A = W.eigen_qcqp_data(n)
    x = cp.Variable(n)
    x.value = np.ones(n)
    prob = cp.Problem(cp.Maximize(cp.quad_form(x, A, assume_PSD=True)), [cp.sum_squares(x) == 1])
    return evaluate(prob, [(None, 11, 0.7)])


def c3_medium():
    m, n = 20000, 1024
    At, x0 = W.logistic_data(m, n, 16)
    x = cp.Variable(n)
    x.value = x0
    obj = cp.sum(cp.logistic(sp.csr_matrix(At) @ x)) + 0.1 * cp.sum(cp.log(1 + cp.power(x, 2))) \
        + 0.01 * cp.sum(cp.exp(-x))
    return evaluate(cp.Problem(cp.Minimize(obj)), [(None, 21, 1.0)])


def c4_medium():
    n, k = 512, 8
    P, q, rng = W.qcqp_data(n, k)
    X = rng.uniform(-1, 1, (4096, n))                    # the benchmark's start points (same rng stream)
    x = cp.Variable(n, bounds=[-1, 1])
    x.value = X[0]
    cons = [cp.quad_form(x, P[i], assume_PSD=True) + q[i] @ x <= 1 for i in range(1, k + 1)]
    prob = cp.Problem(cp.Minimize(cp.quad_form(x, P[0], assume_PSD=True) + q[0] @ x), cons)
    picks = [0, 1000, 2048, 4095]
    out = evaluate(prob, [(X[b], 400 + b, 1.0 if b % 2 == 0 else 0.25) for b in picks])
    out["starts"] = np.array(picks, np.int64)
    return out


def c5_medium():
    N, m = 100000, 50000
    A, x0 = W.microbench_data(N, m, 10)
    ops = [cp.exp, cp.logistic, cp.sin, cp.cos, cp.tanh, cp.sinh,
           lambda v: cp.power(v, 2), lambda v: cp.power(v, 3)]           # = dnlp_b200.workloads.C5_OPS
    seg = N // len(ops)
    xs = [cp.Variable(seg) for _ in ops]
    for s, v in enumerate(xs):
        v.value = x0[s * seg:(s + 1) * seg]
    Ac = sp.csc_matrix(A)
    g, f = 0, 0
    for s, (op, v) in enumerate(zip(ops, xs)):
        g = g + sp.csr_matrix(Ac[:, s * seg:(s + 1) * seg]) @ op(v)
        f = f + cp.sum(op(v))
    return evaluate(cp.Problem(cp.Minimize(f), [g == 0]), [(None, 51, 1.0)])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for fn in (c2_medium, c3_medium, c4_medium, c5_medium):
        if only and fn.__name__ not in only:
            continue
        out = fn()
        np.savez_compressed(os.path.join(OUT, fn.__name__ + ".npz"), **out)
        print("%-10s n=%-7d m=%-7d nnzJ=%-8d nnzH=%-8d points=%d" % (
            fn.__name__, out["n"], out["m"], out["nnz_jac"], out["nnz_hess"], out["npoints"]))
