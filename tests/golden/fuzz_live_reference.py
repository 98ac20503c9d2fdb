"""Differential fuzzing against the LIVE reference (build container only; needs /root/reference).

    python tests/golden/fuzz_live_reference.py [first_seed last_seed]

Random raw smooth problems are written with the reference's own atoms (no Dnlp2Smooth, like
tests/golden/make_golden_atoms.py), sent through the reference's ``Bounds`` + ``Oracles`` (nlp_solver.py:81-427) and,
via ``dnlp_b200.frontend_cvxpy.problem_to_ir``, through the DAG compiler (tape semantics: tests/tape_interp.py) and the
oracle port.  Every problem must either be rejected by the reference and by the compiler, or give structures that are
identical entry for entry (order included) and values within rel 1e-10 at two points.  The committed test-suite fuzzes
compiler vs oracle port (tests/test_fuzz_compiler_vs_oracle.py); this script closes the loop to the reference itself.
A summary of the last run is kept in tests/golden/fuzz_live_reference.log.
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

from cvxpy.reductions.solvers.nlp_solvers.nlp_solver import Bounds, Oracles  # noqa: E402

from dnlp_b200.compiler import compile_problem  # noqa: E402
from dnlp_b200.frontend_cvxpy import problem_to_ir  # noqa: E402
from golden_util import assert_close  # noqa: E402
from oracle.dnlp_oracle import RefOracles  # noqa: E402
from tape_interp import TapeInterp  # noqa: E402

MAXDIM = int(os.environ.get("FUZZ_MAXDIM", "4"))       # vectors up to MAXDIM, matrices up to (MAXDIM-1) x (MAXDIM-1)
WRAPS = int(os.environ.get("FUZZ_WRAPS", "3"))         # up to WRAPS-1 affine atoms stacked on every term
UNARY = [cp.exp, cp.log, cp.entr, cp.logistic, cp.sin, cp.cos, cp.tan, cp.sinh, cp.tanh, cp.asinh, cp.atanh, cp.xexp]


def atom(rng, v):
    k = rng.integers(0, len(UNARY) + 3)
    if k < len(UNARY):
        return UNARY[k](v)
    return cp.power(v, [2, 3, 0.5, 1.5, 2.5][rng.integers(0, 5)])


def affine_wrap(rng, e):
    k = rng.integers(0, 20)
    nd, shape = e.ndim, e.shape
    try:
        if k == 0:
            return -e
        if k == 1 and nd >= 1:
            c = rng.uniform(-2, 2, shape)
            c[rng.random(shape) < 0.2] = 0.0
            if not c.any():
                c.flat[0] = 1.0
            return cp.multiply(c, e) if rng.random() < 0.5 else cp.multiply(e, c)
        if k == 2 and nd == 2:
            return cp.sum(e, axis=int(rng.integers(0, 2)), keepdims=bool(rng.integers(0, 2)))
        if k == 3 and nd >= 1:
            return cp.sum(e)
        if k == 4 and nd == 1 and shape[0] >= 3:
            lo = int(rng.integers(0, 2))
            return e[lo::int(rng.integers(1, 3))]
        if k == 5 and nd == 2:
            return e.T
        if k == 6 and nd >= 1:
            A = rng.uniform(-1, 1, (int(rng.integers(1, 4)), shape[0]))
            A[rng.random(A.shape) < 0.3] = 0.0
            if not A.any():
                A[0, 0] = 1.0
            return A @ e
        if k == 7 and nd == 2:
            return e @ rng.uniform(-1, 1, (shape[1], int(rng.integers(1, 4))))
        if k == 8 and nd == 1 and shape[0] >= 2:
            return e[[int(i) for i in rng.integers(0, shape[0], int(rng.integers(1, 4)))]]
        if k == 9 and nd >= 1 and e.size >= 2:
            if nd == 2:
                return cp.reshape(e, (e.size,), order="F")
            for d in (2, 3):
                if e.size % d == 0:
                    return cp.reshape(e, (d, e.size // d), order="F")
        if k == 10 and nd == 0:
            return cp.promote(e, (int(rng.integers(2, 4)),))
        if k == 11 and nd == 2 and 1 in shape:
            m = int(rng.integers(2, 4))
            return cp.broadcast_to(e, (m, shape[1]) if shape[0] == 1 else (shape[0], m))
        if k == 12 and nd == 2:
            r0, c0 = int(rng.integers(0, shape[0])), int(rng.integers(0, shape[1]))
            choice = rng.integers(0, 5)
            if choice == 0:
                return e[r0, :]
            if choice == 1:
                return e[:, c0]
            if choice == 2:
                return e[r0, c0]
            if choice == 3:
                return e[::-1, 0:shape[1]:2]
            return e[r0:, :c0 + 1]
        if k == 13 and nd == 1 and shape[0] >= 2:
            return e[::-1] if rng.random() < 0.5 else e[-1]
        if k == 14 and nd == 1:
            return cp.reshape(e, (shape[0], 1), order="F").T
        if k == 15 and nd == 2:
            A = rng.uniform(-1, 1, (int(rng.integers(1, 3)), shape[0]))
            B = rng.uniform(-1, 1, (shape[1], int(rng.integers(1, 3))))
            return A @ e @ B
        if k == 16 and nd >= 1:
            return float(rng.uniform(-2, 2)) * e + rng.uniform(-1, 1, shape)
        if k == 17 and nd == 2:
            if rng.random() < 0.5:
                mask = rng.random(shape[0]) < 0.6
                if not mask.any():
                    mask[0] = True
                return e[[int(i) for i in np.where(mask)[0]], :]
            return e[[int(i) for i in rng.integers(0, shape[0], 2)], [int(i) for i in rng.integers(0, shape[1], 2)]]
        if k == 18 and nd >= 1:
            return e + e
        if k == 19 and nd == 2:
            return cp.vec(e, order="F")
    except Exception:
        return e
    return e


def random_problem(seed):
    rng = np.random.default_rng(seed)
    shapes = [(int(rng.integers(2, MAXDIM + 1)),), (int(rng.integers(2, MAXDIM)), int(rng.integers(2, MAXDIM))), ()]
    nvars = int(rng.integers(1, 4))
    variables = []
    for i in range(nvars):
        shp = shapes[int(rng.integers(0, 3))]
        v = cp.Variable(shp, name="v%d" % i)
        v.value = rng.uniform(0.3, 0.9, shp if shp != () else None)
        variables.append(v)

    def term():
        v = variables[int(rng.integers(0, nvars))]
        r = rng.random()
        e = None
        if r < 0.18 and nvars >= 2:
            a, b = rng.choice(nvars, 2, replace=False)
            va, vb = variables[a], variables[b]
            if va.shape == vb.shape and va.ndim >= 1:
                e = cp.multiply(va, vb) if rng.random() < 0.6 else cp.rel_entr(va, vb)
            elif va.ndim == 2 and vb.ndim == 2 and va.shape[1] == vb.shape[0]:
                e = va @ vb
            elif va.ndim >= 1 and vb.ndim == 0:
                e = [cp.multiply(vb, va), cp.multiply(va, vb), cp.rel_entr(va, vb), cp.rel_entr(vb, va),
                     cp.quad_over_lin(va, vb)][int(rng.integers(0, 5))]
        elif r < 0.25 and v.ndim >= 1:
            e = v if rng.random() < 0.5 else cp.multiply(rng.uniform(-1, 1, v.shape), v)
        elif r < 0.31 and v.ndim == 1:
            Q = rng.uniform(-1, 1, (v.size, v.size))
            Q = Q + Q.T
            Q[rng.random(Q.shape) < 0.2] = 0.0
            Q = (Q + Q.T) / 2
            e = cp.quad_form(v, Q, assume_PSD=True)
        elif r < 0.35:
            inner = affine_wrap(rng, v)              # an atom of a non-variable: both sides must reject it
            e = UNARY[int(rng.integers(0, len(UNARY)))](inner)
        if e is None:
            e = atom(rng, v)
        for _ in range(int(rng.integers(0, WRAPS))):
            e = affine_wrap(rng, e)
        return e

    def scalarise(e):
        return e if e.ndim == 0 else cp.sum(e)
    obj = scalarise(term())
    for _ in range(int(rng.integers(0, 3))):
        obj = obj + scalarise(term())
    cons = []
    for _ in range(int(rng.integers(0, 4))):
        e = term()
        if rng.random() < 0.4:
            t2 = term()
            if t2.shape == e.shape:
                e = e - t2
        cons.append(e == 0)
    return cp.Problem(cp.Minimize(obj), cons), rng


def outcome(fn):
    try:
        return "ok", fn()
    except Exception as e:            # noqa: BLE001
        return type(e).__name__, None


def check(seed):
    """True: accepted and identical; False: rejected on both sides (or the reference crashes while evaluating);
    None: the reference crashes with a TypeError / IndexError of its own at the structure pass."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob, rng = random_problem(seed)
        bounds = Bounds(prob)
        ref = Oracles(bounds.new_problem, bounds.x0, len(bounds.cl))

        def ref_structures():
            return ref.jacobianstructure(), ref.hessianstructure()
        s_ref, structs = outcome(ref_structures)
        s_ir, pir = outcome(lambda: problem_to_ir(bounds.new_problem, bounds.cl, bounds.cu, bounds.lb, bounds.ub, bounds.x0))
        if s_ir != "ok":
            assert s_ref != "ok", "seed %d: the frontend rejects (%s) what the reference accepts" % (seed, s_ir)
            return False
        s_cmp, tape = outcome(lambda: compile_problem(pir))
        port = RefOracles(pir)
        s_port, _ = outcome(lambda: (port.jacobianstructure(), port.hessianstructure()))
        if s_ref != "ok" and s_cmp != "ok" and s_port != "ok":
            return False                    # rejected by the reference, the compiler and the port
        if s_ref == "IndexError" and s_port == "IndexError":
            # the same family as nlp_solver.py:232: a Jacobian block that comes out EMPTY is an empty FLOAT array in the
            # reference, and the next rule that indexes with it (binary_operators.py:562, index.py) crashes
            return None
        if s_ref == "TypeError":
            # not a rule rejection but a crash inside the reference: quad_over_lin with a 0-d Variable as denominator hands
            # back a bare numpy.float64 where the callers extend() a list (add_expr.py:164, nlp_solver.py:354).  There is
            # nothing to pin; the compiler evaluates these.  Counted separately.
            return None
        if s_ref != "ok":
            assert s_cmp != "ok" and s_port != "ok", "seed %d: reference %s, compiler %s, port %s" % (seed, s_ref, s_cmp, s_port)
            return False
        x0 = np.asarray(bounds.x0, dtype=np.float64)
        if s_cmp != "ok":
            s_grad, _ = outcome(lambda: ref.gradient(x0))
            assert s_grad != "ok", "seed %d: compiler rejects (%s) what the reference evaluates" % (seed, s_cmp)
            return False
        (jr, jc), (hr, hc) = structs
        for got_r, got_c, want_r, want_c, what in ((tape.jac_rows, tape.jac_cols, jr, jc, "jacobian"),
                                                   (tape.hess_rows, tape.hess_cols, hr, hc, "hessian")):
            np.testing.assert_array_equal(got_r, np.asarray(want_r), err_msg="seed %d %s rows" % (seed, what))
            np.testing.assert_array_equal(got_c, np.asarray(want_c), err_msg="seed %d %s cols" % (seed, what))
        pj, ph = port.jacobianstructure(), port.hessianstructure()
        np.testing.assert_array_equal(pj[0], np.asarray(jr)), np.testing.assert_array_equal(pj[1], np.asarray(jc))
        np.testing.assert_array_equal(ph[0], np.asarray(hr)), np.testing.assert_array_equal(ph[1], np.asarray(hc))
        it = TapeInterp(tape)
        m = len(bounds.cl)
        with np.errstate(all="ignore"):
            for _ in range(2):
                x = np.clip(x0 * (1 + 0.1 * rng.standard_normal(x0.size)), 0.05, 0.95)
                lam = rng.standard_normal(m)
                sigma = float(rng.uniform(0.5, 1.5))
                try:
                    want = {"f": ref.objective(x), "grad": np.array(ref.gradient(x), dtype=np.float64),
                            "g": ref.constraints(x) if m else np.zeros(0), "jac": np.array(ref.jacobian(x), dtype=np.float64),
                            "hess": np.array(ref.hessian(x, lam, sigma), dtype=np.float64)}
                except (IndexError, ValueError):
                    return False            # the reference's own evaluation crashes (e.g. nlp_solver.py:232 on an empty block)
                for name in ("f", "grad", "g", "jac", "hess"):
                    got = it.eval(name, x, lam, sigma) if name == "hess" else it.eval(name, x)
                    assert_close(got, want[name], "%s seed %d (compiler)" % (name, seed))
                assert_close(port.objective(x), want["f"], "f seed %d (port)" % seed)
                assert_close(port.gradient(x), want["grad"], "grad seed %d (port)" % seed)
                if m:
                    assert_close(port.constraints(x), want["g"], "g seed %d (port)" % seed)
                assert_close(port.jacobian(x), want["jac"], "jac seed %d (port)" % seed)
                assert_close(port.hessian(x, lam, sigma), want["hess"], "hess seed %d (port)" % seed)
    return True


if __name__ == "__main__":
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 2000)
    accepted = rejected = failed = crashed = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            res = check(seed)
            if res is None:
                crashed += 1
            elif res:
                accepted += 1
            else:
                rejected += 1
        except Exception as e:            # noqa: BLE001
            failed += 1
            print("SEED %d FAILED: %s: %s" % (seed, type(e).__name__, str(e)[:500].replace("\n", " | ")), flush=True)
    print("live-reference fuzz, seeds %d..%d: %d accepted and identical (structures bit-exact, values rel 1e-10), %d rejected "
          "by both sides, %d where the reference itself crashes (TypeError / IndexError of its own), %d FAILURES, %.0f s"
          % (lo, hi, accepted, rejected, crashed, failed, time.time() - t0))
