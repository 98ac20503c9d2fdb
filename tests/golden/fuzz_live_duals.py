"""Dual recovery against first principles, through the LIVE reference's chain (build container only).

    python tests/golden/fuzz_live_duals.py [first_seed last_seed]

The reference hands back no duals (ipopt_nlpif.py:100); ``dnlp_b200.nlp_solver.install_dual_recovery()`` maps the
solver's ``mult_g`` through the reference's own invert chain (reductions/canonicalization.py:76-84) onto
``constraint.dual_value``.  For random equality-constrained problems (convex objectives of smooth atoms of affine
expressions, linear and nonlinear equalities, Minimize and Maximize) the smooth problem the chain produces is solved by
the Newton-KKT stand-in for IPOPT (tests/kkt_newton.py) on the reference's own ``Oracles``; the recovered duals must make
the Lagrangian of the ORIGINAL problem, in the original variables, stationary:
    grad f(x) + sum_i J_i(x)' dual_i = 0      (f: the expression being MINIMISED, i.e. -objective for a Maximize
                                                problem - cvxpy's FlipObjective does not touch the duals; J_i: Jacobian
                                                of lhs_i - rhs_i)
with every gradient taken by central differences of the cvxpy expressions' own values - nothing of this repo.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import cp  # noqa: E402  (loads the reference)

from cvxpy.reductions.cvx_attr2constr import CvxAttr2Constr  # noqa: E402
from cvxpy.reductions.dnlp2smooth.dnlp2smooth import Dnlp2Smooth  # noqa: E402
from cvxpy.reductions.flip_objective import FlipObjective  # noqa: E402
from cvxpy.reductions.solvers.nlp_solvers.ipopt_nlpif import IPOPT  # noqa: E402
from cvxpy.reductions.solvers.solving_chain import SolvingChain  # noqa: E402

import dnlp_b200.nlp_solver as gpu  # noqa: E402
import kkt_newton  # noqa: E402


def random_problem(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 7))
    x = cp.Variable(n, name="x")
    x.value = rng.uniform(0.2, 0.6, n)
    convex = [lambda e: cp.sum(cp.exp(e)), lambda e: cp.sum(cp.logistic(e)), lambda e: cp.sum_squares(e),
              lambda e: cp.sum(cp.power(e, 4)), lambda e: -cp.sum(cp.entr(e + 1.5)), lambda e: cp.sum(cp.cosh(e)) if hasattr(cp, "cosh") else cp.sum_squares(e)]
    obj = cp.sum_squares(x - rng.uniform(-0.5, 0.5, n))                  # keeps the problem strictly convex
    for _ in range(int(rng.integers(1, 3))):
        A = rng.uniform(-1, 1, (int(rng.integers(1, 4)), n))
        obj = obj + float(rng.uniform(0.2, 1.0)) * convex[int(rng.integers(0, 5))](A @ x + rng.uniform(-0.2, 0.2, A.shape[0]))
    cons = []
    k = int(rng.integers(1, min(n - 1, 3) + 1))
    A = rng.standard_normal((k, n))
    cons.append(A @ x == A @ x.value + rng.uniform(-0.1, 0.1, k))
    if rng.random() < 0.5:                                               # a smooth nonlinear equality
        cons.append(cp.sum(cp.exp(x)) == float(np.exp(x.value).sum() * rng.uniform(0.95, 1.05)))
    maximize = rng.random() < 0.3
    prob = cp.Problem(cp.Maximize(-obj) if maximize else cp.Minimize(obj), cons)
    return prob, x, obj, cons


def num_grad(expr, x, xv, h=1e-6):
    out = []
    for j in range(xv.size):
        e = np.zeros_like(xv)
        e[j] = h
        x.value = xv + e
        up = np.atleast_1d(np.asarray(expr.value, dtype=np.float64)).reshape(-1).copy()
        x.value = xv - e
        dn = np.atleast_1d(np.asarray(expr.value, dtype=np.float64)).reshape(-1).copy()
        out.append((up - dn) / (2 * h))
    x.value = xv
    return np.stack(out, axis=1)             # (size of expr, n)


def check(seed):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob, x, obj, cons = random_problem(seed)
        if not prob.is_dnlp():
            return False
        maximize = type(prob.objective) == cp.Maximize
        chain = SolvingChain(reductions=([FlipObjective()] if maximize else []) + [CvxAttr2Constr(reduce_bounds=False), Dnlp2Smooth(), IPOPT()])
        gpu.install_dual_recovery()
        try:
            data, inverse_data = chain.apply(problem=prob)
            o = data["oracles"]
            if np.any(np.asarray(data["cl"]) != np.asarray(data["cu"])) or np.any(np.isfinite(data["lb"])) or np.any(np.isfinite(data["ub"])):
                return False                 # the stand-in solves equality-constrained problems only
            try:
                xs, lam, f, iters = kkt_newton.solve(o, data["x0"], tol=1e-11, max_iter=80)
            except np.linalg.LinAlgError:
                return False
            r = np.asarray(o.constraints(xs), dtype=np.float64)
            if iters >= 80 or not np.all(np.isfinite(xs)) or np.abs(r).max() > 1e-8:
                return False                 # Newton without globalisation did not get there: nothing to check
            prob.unpack_results({"status": 0, "x": xs, "obj_val": f, "mult_g": lam, "iterations": iters}, chain, inverse_data)
        finally:
            gpu.uninstall()
        xv = np.asarray(x.value, dtype=np.float64).copy()
        resid = num_grad(obj, x, xv)[0]                                   # the minimised function, whatever the sense
        for c in cons:
            dual = np.atleast_1d(np.asarray(c.dual_value, dtype=np.float64)).reshape(-1)
            assert dual.size == c.size, "seed %d: dual of %s has %d entries" % (seed, c, dual.size)
            J = num_grad(c.args[0] - c.args[1], x, xv)
            # FlipObjective.invert (reductions/flip_objective.py) only negates the optimal value: the duals of a Maximize
            # problem are those of the minimisation it was turned into, so the sign is the same in both senses
            resid = resid + J.T @ dual
        scale = max(1.0, float(np.abs(num_grad(obj, x, xv)[0]).max()))
        assert np.abs(resid).max() < 2e-5 * scale, "seed %d: Lagrangian not stationary, residual %s" % (seed, resid)
        for c in cons:                                                    # and primal feasibility in the original variables
            assert np.abs(np.asarray((c.args[0] - c.args[1]).value)).max() < 1e-7
    return True


if __name__ == "__main__":
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 500)
    ok = skipped = failed = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            if check(seed):
                ok += 1
            else:
                skipped += 1
        except Exception as e:          # noqa: BLE001
            failed += 1
            print("SEED %d FAILED: %s: %s" % (seed, type(e).__name__, str(e)[:400].replace("\n", " | ")), flush=True)
    print("dual recovery fuzz, seeds %d..%d: %d problems whose recovered duals make the ORIGINAL Lagrangian stationary "
          "(central differences of the cvxpy expressions), %d skipped (stand-in solver did not converge / not equality "
          "form), %d FAILURES, %.0f s" % (lo, hi, ok, skipped, failed, time.time() - t0))
