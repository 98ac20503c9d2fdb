"""Differential fuzzing against the LIVE reference THROUGH ITS REDUCTION CHAIN (build container only).

    python tests/golden/fuzz_live_chain.py [first_seed last_seed]

Where tests/golden/fuzz_live_reference.py feeds raw rules, this script writes random DNLP problems the way a user
would - atoms of affine expressions, nonsmooth atoms (abs, maximum, minimum, norm1, norm2, huber), bounds, equality and
inequality constraints - and lets the reference canonicalise them: ``FlipObjective`` / ``CvxAttr2Constr`` /
``Dnlp2Smooth`` / ``IPOPT.apply`` (tests/golden/make_golden.py:reference_data).  The smooth problem then goes through
``frontend_cvxpy.data_to_ir`` into the compiler and the oracle port.  Checked per accepted problem: x0, bounds and
constraint bounds of the IR equal the chain's, structures identical with their order, the five outputs within rel
1e-10 (constraint rows that cancel: atol 1e-9) at two points.  Summary of the last run: fuzz_live_chain.log.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_golden as mg  # noqa: E402  (loads the reference)
from make_golden import cp  # noqa: E402

from dnlp_b200.compiler import compile_problem  # noqa: E402
from dnlp_b200.frontend_cvxpy import data_to_ir  # noqa: E402
from golden_util import assert_close  # noqa: E402
from oracle.dnlp_oracle import RefOracles  # noqa: E402
from tape_interp import TapeInterp  # noqa: E402

SMOOTH = [cp.exp, cp.log, cp.entr, cp.logistic, cp.sin, cp.cos, cp.tan, cp.sinh, cp.tanh, cp.asinh, cp.atanh, cp.square,
          cp.sqrt, lambda e: cp.power(e, 3), lambda e: cp.power(e, 1.5), lambda e: cp.power(e, 0.5)]
NONSMOOTH = [cp.abs, cp.pos, cp.neg, lambda e: cp.huber(e, 0.5), lambda e: cp.maximum(e, 0.3), lambda e: cp.minimum(e, 0.7)]


def affine(rng, v):
    k = rng.integers(0, 6)
    if k == 0 or v.ndim == 0:
        return float(rng.uniform(0.5, 1.5)) * v + float(rng.uniform(0.0, 0.2))
    if k == 1 and v.ndim == 1:
        A = rng.uniform(0.1, 1.0, (int(rng.integers(1, 4)), v.shape[0]))
        return A @ v + rng.uniform(0.0, 0.2, A.shape[0])
    if k == 2 and v.ndim == 2:
        return v.T
    if k == 3 and v.ndim == 1 and v.shape[0] >= 3:
        return v[1:]
    if k == 4 and v.ndim >= 1:
        return cp.multiply(rng.uniform(0.5, 1.5, v.shape), v)
    return v


def term(rng, variables):
    v = variables[int(rng.integers(0, len(variables)))]
    r = rng.random()
    arg = affine(rng, v) if rng.random() < 0.6 else v
    if r < 0.55:
        e = SMOOTH[int(rng.integers(0, len(SMOOTH)))](arg)
    elif r < 0.75:
        e = NONSMOOTH[int(rng.integers(0, len(NONSMOOTH)))](arg)
    elif r < 0.82 and arg.ndim == 1:
        e = [cp.norm1, cp.norm2, cp.sum_squares][int(rng.integers(0, 3))](arg)
    elif r < 0.90 and len(variables) >= 2:
        a, b = rng.choice(len(variables), 2, replace=False)
        va, vb = variables[a], variables[b]
        if va.shape == vb.shape and va.ndim >= 1:
            e = [cp.multiply(va, vb), cp.rel_entr(va, vb), cp.kl_div(va, vb)][int(rng.integers(0, 3))]
        elif va.ndim == 2 and vb.ndim == 2 and va.shape[1] == vb.shape[0]:
            e = va @ vb
        elif va.ndim == 1 and vb.ndim == 0:
            e = cp.quad_over_lin(va, vb)
        else:
            e = cp.exp(arg)
    elif arg.ndim == 1:
        Q = rng.uniform(-1, 1, (arg.shape[0], arg.shape[0]))
        Q = Q @ Q.T + 0.1 * np.eye(arg.shape[0])
        e = cp.quad_form(arg, Q, assume_PSD=True)
    else:
        e = cp.exp(arg)
    if rng.random() < 0.3 and e.ndim >= 1:
        e = cp.multiply(rng.uniform(0.2, 2.0, e.shape), e)
    return e


def random_problem(seed):
    rng = np.random.default_rng(seed)
    shapes = [(int(rng.integers(2, 5)),), (int(rng.integers(2, 4)), int(rng.integers(2, 4))), ()]
    variables = []
    for i in range(int(rng.integers(1, 4))):
        shp = shapes[int(rng.integers(0, 3))]
        kw = {}
        u = rng.random()
        if u < 0.2:
            kw["nonneg"] = True
        elif u < 0.4:
            kw["bounds"] = [0.1, 2.0]
        v = cp.Variable(shp, name="v%d" % i, **kw)
        if rng.random() < 0.7:
            v.value = rng.uniform(0.3, 0.9, shp if shp != () else None)
        variables.append(v)
    obj = 0
    for _ in range(int(rng.integers(1, 4))):
        e = term(rng, variables)
        obj = obj + (e if e.ndim == 0 else cp.sum(e))
    cons = []
    for _ in range(int(rng.integers(0, 4))):
        e = term(rng, variables)
        u = rng.random()
        rhs = float(rng.uniform(0.5, 2.0))
        cons.append(e == rhs if u < 0.4 else (e <= rhs if u < 0.7 else e >= rhs * 0.1))
    for v in variables:
        if rng.random() < 0.3 and v.ndim >= 1:
            cons.append(v >= 0.05)
    sense = cp.Minimize if rng.random() < 0.7 else cp.Maximize
    return cp.Problem(sense(obj), cons), rng


def check(seed):
    """True: accepted and identical; False: not DNLP / rejected by the chain or by every implementation."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob, rng = random_problem(seed)
        try:
            if not prob.is_dnlp():
                return False
            data = mg.reference_data(prob)
        except Exception:
            return False
        ref = data["oracles"]
        try:
            jr, jc = ref.jacobianstructure()
            hr, hc = ref.hessianstructure()
            s_ref = "ok"
        except Exception as e:          # noqa: BLE001
            s_ref = type(e).__name__
        try:
            pir = data_to_ir(data)
            tape = compile_problem(pir)
            s_cmp = "ok"
        except Exception as e:          # noqa: BLE001
            s_cmp = type(e).__name__
        if s_ref != "ok" or s_cmp != "ok":
            assert s_ref != "ok" and s_cmp != "ok", "seed %d: reference %s vs compiler %s" % (seed, s_ref, s_cmp)
            return False
        np.testing.assert_array_equal(pir.x0, np.asarray(data["x0"], dtype=np.float64), err_msg="seed %d x0" % seed)
        for key in ("lb", "ub", "cl", "cu"):
            np.testing.assert_array_equal(getattr(pir, key), np.asarray(data[key], dtype=np.float64), err_msg="seed %d %s" % (seed, key))
        port = RefOracles(pir)
        for (gr, gc), (wr, wc), what in (((tape.jac_rows, tape.jac_cols), (jr, jc), "jacobian"),
                                         ((tape.hess_rows, tape.hess_cols), (hr, hc), "hessian"),
                                         (port.jacobianstructure(), (jr, jc), "jacobian (port)"),
                                         (port.hessianstructure(), (hr, hc), "hessian (port)")):
            np.testing.assert_array_equal(gr, np.asarray(wr), err_msg="seed %d %s rows" % (seed, what))
            np.testing.assert_array_equal(gc, np.asarray(wc), err_msg="seed %d %s cols" % (seed, what))
        it = TapeInterp(tape)
        x0 = np.asarray(data["x0"], dtype=np.float64)
        m = len(data["cl"])
        lo = np.where(np.isfinite(data["lb"]), data["lb"], -np.inf)
        hi = np.where(np.isfinite(data["ub"]), data["ub"], np.inf)
        with np.errstate(all="ignore"):
            for _ in range(2):
                x = np.clip(x0 + 0.05 * rng.standard_normal(x0.size), np.maximum(lo, x0 - 0.2), np.minimum(hi, x0 + 0.2))
                lam = rng.standard_normal(m)
                sigma = float(rng.uniform(0.5, 1.5))
                try:
                    want = {"f": ref.objective(x), "grad": np.array(ref.gradient(x), dtype=np.float64),
                            "g": np.asarray(ref.constraints(x), dtype=np.float64) if m else np.zeros(0),
                            "jac": np.array(ref.jacobian(x), dtype=np.float64),
                            "hess": np.array(ref.hessian(x, lam, sigma), dtype=np.float64)}
                except (IndexError, ValueError, TypeError):
                    return False        # the reference's own evaluation crashes
                for name in ("f", "grad", "g", "jac", "hess"):
                    got = it.eval(name, x, lam, sigma) if name == "hess" else it.eval(name, x)
                    assert_close(got, want[name], "%s seed %d (compiler)" % (name, seed), atol=1e-9 if name == "g" else 1e-12)
                assert_close(port.objective(x), want["f"], "f seed %d (port)" % seed)
                assert_close(port.gradient(x), want["grad"], "grad seed %d (port)" % seed)
                if m:
                    assert_close(port.constraints(x), want["g"], "g seed %d (port)" % seed, atol=1e-9)
                assert_close(port.jacobian(x), want["jac"], "jac seed %d (port)" % seed)
                assert_close(port.hessian(x, lam, sigma), want["hess"], "hess seed %d (port)" % seed)
    return True


if __name__ == "__main__":
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 1000)
    accepted = skipped = failed = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            if check(seed):
                accepted += 1
            else:
                skipped += 1
        except Exception as e:          # noqa: BLE001
            failed += 1
            print("SEED %d FAILED: %s: %s" % (seed, type(e).__name__, str(e)[:500].replace("\n", " | ")), flush=True)
    print("live-reference chain fuzz, seeds %d..%d: %d DNLP problems canonicalised by the reference and identical in the "
          "compiler and the port (x0 / bounds equal, structures bit-exact, values rel 1e-10), %d not DNLP or rejected by "
          "all, %d FAILURES, %.0f s" % (lo, hi, accepted, skipped, failed, time.time() - t0))
