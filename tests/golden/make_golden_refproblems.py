"""Golden vectors for the PROBLEMS of the reference's own NLP test-suite (build container only).

    python tests/golden/make_golden_refproblems.py      # writes tests/golden/refproblems/*.npz

The reference's problem-level tests (cvxpy/tests/NLP_tests/test_*.py) build a problem and call
``prob.solve(nlp=True, solver=IPOPT | KNITRO, ...)``; without a solver in the image they are skipped.  This script
imports those test modules from where they lie and calls every test function with ``Problem.solve`` wrapped: an
``nlp=True`` call records the problem and stops the test (there is no solution to assert on).  Every distinct
problem then goes through tests/golden/make_golden.py's pipeline - the reference's own reduction chain
(``FlipObjective`` / ``CvxAttr2Constr`` / ``Dnlp2Smooth`` / ``IPOPT.apply``) and its ``Oracles`` at several points.

Nothing is copied from the reference's tests: they are imported and executed.
"""
import glob
import hashlib
import importlib.util
import inspect
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference)
from make_golden import cp  # noqa: E402

REF_TESTS = "/root/reference/cvxpy/tests/NLP_tests"
OUT = os.path.join(HERE, "refproblems")
MAX_N, MAX_NNZ = 10000, 300000
CAPTURED = []            # (test id, problem)
_current = [""]


class _Captured(Exception):
    pass


def _patch_solve():
    original = cp.Problem.solve

    def solve(self, *args, **kwargs):
        if kwargs.get("nlp"):
            CAPTURED.append((_current[0], self))
            raise _Captured()
        try:                                # a convex solve some tests run first, for comparison
            return original(self, *args, **kwargs)
        except Exception:                   # no conic solver in this image: carry on to the nlp=True solve
            return np.nan
    cp.Problem.solve = solve
    return original


def run_reference_tests():
    ran = 0
    for path in sorted(glob.glob(os.path.join(REF_TESTS, "test_*.py"))):
        mod_name = "_refp_" + os.path.splitext(os.path.basename(path))[0]
        spec = importlib.util.spec_from_file_location(mod_name, path)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception as e:
            print("  (module %s not importable: %s: %s)" % (os.path.basename(path), type(e).__name__, str(e)[:80]))
            continue
        tests = []
        for name, obj in inspect.getmembers(mod):
            if inspect.isclass(obj) and name.startswith("Test") and obj.__module__ == mod_name:
                try:
                    inst = obj()
                except Exception:
                    continue
                for hook in ("setUp", "setup_method"):
                    if hasattr(inst, hook):
                        try:
                            getattr(inst, hook)() if hook == "setUp" else getattr(inst, hook)(None)
                        except Exception:
                            pass
                tests += [("%s::%s::%s" % (os.path.basename(path), name, m), getattr(inst, m))
                          for m in sorted(dir(inst)) if m.startswith("test_")]
            elif inspect.isfunction(obj) and name.startswith("test_") and obj.__module__ == mod_name:
                tests.append(("%s::%s" % (os.path.basename(path), name), obj))
        for tid, fn in tests:
            _current[0] = tid
            try:
                params = [p for p in inspect.signature(fn).parameters.values() if p.default is inspect.Parameter.empty]
            except (TypeError, ValueError):
                params = []
            kwargs, ok = {}, True
            for prm in params:              # pytest would fill these in: the solver name, or a module-level fixture
                if prm.name == "solver":
                    kwargs["solver"] = "IPOPT"
                    continue
                fx = getattr(mod, prm.name, None)
                made = False
                for getter in (lambda f: f.__wrapped__(), lambda f: f._get_wrapped_function()(), lambda f: f()):
                    try:
                        kwargs[prm.name] = getter(fx)
                        made = True
                        break
                    except Exception:
                        pass
                ok = ok and made
            if not ok:
                print("  (cannot supply the arguments of %s)" % tid)
                continue
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    fn(**kwargs)
            except _Captured:
                pass
            except Exception:
                pass                        # the rest of the test needs a solution; what was captured stays
            ran += 1
    return ran


def main():
    original = _patch_solve()
    ran = run_reference_tests()
    cp.Problem.solve = original
    print("reference problem tests executed: %d; nlp=True solves captured: %d" % (ran, len(CAPTURED)))
    os.makedirs(OUT, exist_ok=True)
    for old in glob.glob(os.path.join(OUT, "*.npz")):
        os.remove(old)
    seen, index, k = set(), [], 0
    for tid, prob in CAPTURED:
        name = "refp_%03d" % k
        try:
            n, m, nj, nh = mg.make_one(name, lambda p=prob: p, OUT)
        except Exception as e:              # the reference itself rejects the problem (not DNLP, unsupported atom, ...)
            index.append("-\t%s\tREFERENCE ERROR %s: %s" % (tid, type(e).__name__, str(e)[:100].replace("\n", " ")))
            print(index[-1])
            continue
        path = os.path.join(OUT, name + ".npz")
        z = np.load(path)
        sig = hashlib.sha256(str(z["ir_json"]).encode())
        for key in sorted(f for f in z.files if f.startswith("ir_a")):
            sig.update(np.ascontiguousarray(z[key]).tobytes())
        sig = sig.hexdigest()
        if sig in seen or n > MAX_N or max(nj, nh) > MAX_NNZ:
            os.remove(path)
            why = "duplicate of an earlier fixture" if sig in seen else "too large for a committed fixture"
            index.append("-\t%s\t%s (n=%d m=%d nnzJ=%d nnzH=%d)" % (tid, why, n, m, nj, nh))
            print(index[-1])
            continue
        seen.add(sig)
        k += 1
        index.append("%s\t%s\tn=%d m=%d nnzJ=%d nnzH=%d" % (name, tid, n, m, nj, nh))
        print(index[-1])
    with open(os.path.join(OUT, "INDEX.tsv"), "w") as f:
        f.write("fixture\treference test whose first nlp=True solve it is\tsizes / why there is no fixture\n")
        f.write("\n".join(index) + "\n")
    print("fixtures written: %d" % k)


if __name__ == "__main__":
    main()
