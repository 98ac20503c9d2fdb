"""Differential fuzzing of PARAMETER SLOTS against the live reference (build container only).

    python tests/golden/fuzz_live_parameters.py [first_seed last_seed]

Random DNLP problems with ``cp.Parameter`` objects where a user puts them - added constants, elementwise and scalar
factors, right-hand sides of constraints - go through the reference's reduction chain.  The tape is compiled ONCE at
the first parameter setting; for two further settings only the slot values change (``TapeInterp.set_params``, what
``GpuOracles.set_parameters`` / ``dnlp_set_params`` do on the device), the reference re-reads ``Parameter.value``
inside its rules (expressions/constants/parameter.py:35).  Checked per setting: the fingerprint is unchanged (no
recompile), structures identical, the five outputs within rel 1e-10.  Summary of the last run:
fuzz_live_parameters.log.
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

from dnlp_b200.compile_cache import fingerprint  # noqa: E402
from dnlp_b200.compiler import compile_problem  # noqa: E402
from dnlp_b200.frontend_cvxpy import data_to_ir  # noqa: E402
from golden_util import assert_close  # noqa: E402
from tape_interp import TapeInterp  # noqa: E402

FROZEN = bool(int(os.environ.get("FUZZ_FROZEN", "0")))   # also put parameters where they cannot be slots (matrix of a product)
SMOOTH = [cp.exp, cp.log, cp.entr, cp.logistic, cp.sin, cp.cos, cp.tanh, cp.square, cp.sqrt, lambda e: cp.power(e, 3)]


def random_problem(seed):
    rng = np.random.default_rng(seed)
    shapes = [(int(rng.integers(2, 5)),), (int(rng.integers(2, 4)), int(rng.integers(2, 4))), ()]
    variables, params = [], []
    for i in range(int(rng.integers(1, 4))):
        shp = shapes[int(rng.integers(0, 3))]
        v = cp.Variable(shp, name="v%d" % i, **({"bounds": [0.1, 2.0]} if rng.random() < 0.3 else {}))
        v.value = rng.uniform(0.3, 0.9, shp if shp != () else None)
        variables.append(v)

    def new_param(shape, lo=0.5, hi=1.5):
        p = cp.Parameter(shape, name="p%d" % len(params))
        p.value = np.asarray(rng.uniform(lo, hi, shape if shape != () else None), dtype=np.float64)
        params.append((p, lo, hi))
        return p

    def term():
        v = variables[int(rng.integers(0, len(variables)))]
        arg = v
        u = rng.random()
        if u < 0.25:                                   # parameter added inside the atom's (affine) argument
            arg = v + new_param(v.shape, 0.0, 0.3)
        elif u < 0.45 and v.ndim >= 1:                 # elementwise parameter factor inside
            arg = cp.multiply(new_param(v.shape), v)
        elif u < 0.55:
            arg = new_param(()) * v
        elif u < 0.65 and v.ndim == 1 and FROZEN:      # COEFFICIENT position: the frontend freezes the value into the tape
            arg = new_param((int(rng.integers(1, 4)), v.shape[0])) @ v
        if FROZEN and v.ndim == 1 and rng.random() < 0.1:
            P = new_param((v.shape[0], v.shape[0]), 0.1, 1.0)
            if rng.random() < 0.5:
                return cp.quad_form(v, P, assume_PSD=True)          # the matrix of a quad_form
            return cp.sum(P @ v)
        e = SMOOTH[int(rng.integers(0, len(SMOOTH)))](arg)
        w = rng.random()
        if w < 0.25 and e.ndim >= 1:                   # ... and outside
            e = cp.multiply(new_param(e.shape), e)
        elif w < 0.4:
            e = new_param(()) * e
        elif w < 0.55:
            e = e + new_param(e.shape, -0.5, 0.5)
        return e
    obj = 0
    for _ in range(int(rng.integers(1, 4))):
        e = term()
        obj = obj + (e if e.ndim == 0 else cp.sum(e))
    cons = []
    for _ in range(int(rng.integers(0, 4))):
        e = term()
        rhs = new_param(e.shape, 0.5, 2.0) if rng.random() < 0.6 else float(rng.uniform(0.5, 2.0))
        u = rng.random()
        cons.append(e == rhs if u < 0.4 else (e <= rhs if u < 0.7 else e >= rhs))
    return cp.Problem((cp.Minimize if rng.random() < 0.7 else cp.Maximize)(obj), cons), params, rng


def compare(seed, data, tape, it, rng, setting):
    ref = data["oracles"]
    jr, jc = ref.jacobianstructure()
    hr, hc = ref.hessianstructure()
    for got, want, what in ((tape.jac_rows, jr, "jac rows"), (tape.jac_cols, jc, "jac cols"), (tape.hess_rows, hr, "hess rows"),
                            (tape.hess_cols, hc, "hess cols")):
        np.testing.assert_array_equal(got, np.asarray(want), err_msg="seed %d setting %d %s" % (seed, setting, what))
    x0 = np.asarray(data["x0"], dtype=np.float64)
    m = len(data["cl"])
    lo = np.where(np.isfinite(data["lb"]), data["lb"], -np.inf)
    hi = np.where(np.isfinite(data["ub"]), data["ub"], np.inf)
    with np.errstate(all="ignore"):
        for _ in range(2):
            x = np.clip(x0 + 0.05 * rng.standard_normal(x0.size), np.maximum(lo, x0 - 0.2), np.minimum(hi, x0 + 0.2))
            lam = rng.standard_normal(m)
            sigma = float(rng.uniform(0.5, 1.5))
            want = {"f": ref.objective(x), "grad": np.array(ref.gradient(x), dtype=np.float64),
                    "g": np.asarray(ref.constraints(x), dtype=np.float64) if m else np.zeros(0),
                    "jac": np.array(ref.jacobian(x), dtype=np.float64),
                    "hess": np.array(ref.hessian(x, lam, sigma), dtype=np.float64)}
            for name in ("f", "grad", "g", "jac", "hess"):
                got = it.eval(name, x, lam, sigma) if name == "hess" else it.eval(name, x)
                assert_close(got, want[name], "%s seed %d setting %d" % (name, seed, setting), atol=1e-9 if name == "g" else 1e-12)


def check(seed):
    """(accepted, number of parameter slots) or (False, 0) when the problem is not DNLP / rejected / has no parameter."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob, params, rng = random_problem(seed)
        if not params:
            return False, 0
        try:
            if not prob.is_dnlp():
                return False, 0
            data = mg.reference_data(prob)
            data["oracles"].jacobianstructure(), data["oracles"].hessianstructure()
        except Exception:
            return False, 0
        try:
            pir = data_to_ir(data)
            tape = compile_problem(pir)
        except Exception as e:          # noqa: BLE001
            raise AssertionError("seed %d: the reference accepts, the compiler raises %s: %s" % (seed, type(e).__name__, e))
        it = TapeInterp(tape)
        fp = fingerprint(pir)
        try:
            compare(seed, data, tape, it, rng, 0)
        except (IndexError, TypeError):
            return False, 0             # the reference's own evaluation crashes
        for setting in (1, 2):
            for p, lo, hi in params:
                p.value = np.asarray(rng.uniform(lo, hi, p.shape if p.shape != () else None), dtype=np.float64)
            data2 = mg.reference_data(prob)
            pir2 = data_to_ir(data2)
            if FROZEN and fingerprint(pir2) != fp:
                # a parameter in a coefficient position is part of the tape: its new value is a new fingerprint, which
                # is what makes the compile cache recompile instead of serving stale coefficients
                tape2 = compile_problem(pir2)
                compare(seed, data2, tape2, TapeInterp(tape2), rng, setting)
                recompiled[0] += 1
                continue
            assert fingerprint(pir2) == fp, "seed %d: new parameter VALUES changed the fingerprint" % seed
            assert pir2.n_params == tape.n_params
            it.set_params(pir2.param_values())          # no recompile: the same tape, new slot values
            compare(seed, data2, tape, it, rng, setting)
        return True, tape.n_params


recompiled = [0]


if __name__ == "__main__":
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 1000)
    accepted = skipped = failed = slots = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            ok, n = check(seed)
            accepted += bool(ok)
            skipped += not ok
            slots += n
        except Exception as e:          # noqa: BLE001
            failed += 1
            print("SEED %d FAILED: %s: %s" % (seed, type(e).__name__, str(e)[:500].replace("\n", " | ")), flush=True)
    print("live-reference parameter fuzz, seeds %d..%d: %d problems compiled once and identical to the reference at three "
          "parameter settings each (%d parameter slots in total; fingerprint unchanged, structures bit-exact, values rel "
          "1e-10), %d not DNLP / rejected, %d FAILURES, %.0f s" % (lo, hi, accepted, slots, skipped, failed, time.time() - t0))
    if FROZEN:
        print("  FUZZ_FROZEN=1: parameters also in coefficient positions (matrix of a product): %d settings changed the "
              "fingerprint and were recompiled, the others reused the tape" % recompiled[0])
