"""Golden vectors for EVERY expression the reference's own rule tests differentiate (build container only).

    python tests/golden/make_golden_reftests.py      # writes tests/golden/reftests/*.npz

The reference's tier-1 tests (cvxpy/tests/NLP_tests/jacobian_tests/*.py, hess_tests/*.py) build raw expressions
and call ``expr.jacobian()`` / ``expr.hess_vec(vec)`` on them.  This script runs those test functions unmodified,
with ``Atom.jacobian`` / ``Atom.hess_vec`` (atoms/atom.py:501-561) wrapped so that every TOP-LEVEL call records the
expression it was made on, at the variable values the test gave it.  Each distinct expression then goes through the
same pipeline as tests/golden/make_golden_atoms.py: ``Problem(Minimize(0), [expr == 0])`` -> the reference's ``Bounds``
+ ``Oracles`` -> structures (with their ORDER, which the reference's own tests do not pin: they scatter into dense
matrices) and values of all callbacks at two points, or the exception type where the reference rejects the rule.

Nothing is copied from the reference's tests: they are imported from where they lie and executed.
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
import make_golden_atoms as mga  # noqa: E402  (loads the reference)
from make_golden import cp  # noqa: E402

from cvxpy.atoms.atom import Atom  # noqa: E402

REF_TESTS = "/root/reference/cvxpy/tests/NLP_tests"
OUT = os.path.join(HERE, "reftests")
RECORDED = []            # (test id, "jac" | "hess", expression)
_depth = [0]
_current = [""]


def _wrap(method_name, kind):
    original = getattr(Atom, method_name)

    def wrapper(self, *args, **kwargs):
        if _depth[0] == 0:
            RECORDED.append((_current[0], kind, self))
        _depth[0] += 1
        try:
            return original(self, *args, **kwargs)
        finally:
            _depth[0] -= 1
    setattr(Atom, method_name, wrapper)
    return original


def run_reference_tests():
    """Import every jacobian / hess test module of the reference and call its test functions."""
    ran = failed = 0
    for sub in ("jacobian_tests", "hess_tests"):
        for path in sorted(glob.glob(os.path.join(REF_TESTS, sub, "test_*.py"))):
            mod_name = "_ref_%s_%s" % (sub, os.path.splitext(os.path.basename(path))[0])
            spec = importlib.util.spec_from_file_location(mod_name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            tests = []
            for name, obj in inspect.getmembers(mod):
                if inspect.isclass(obj) and name.startswith("Test") and obj.__module__ == mod_name:
                    inst = obj()
                    tests += [("%s::%s::%s" % (os.path.basename(path), name, m), getattr(inst, m))
                              for m in sorted(dir(inst)) if m.startswith("test_")]
                elif inspect.isfunction(obj) and name.startswith("test_") and obj.__module__ == mod_name:
                    tests.append(("%s::%s" % (os.path.basename(path), name), obj))
            for tid, fn in tests:
                _current[0], _depth[0] = tid, 0
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        fn()
                    ran += 1
                except Exception as e:          # a test that needs a fixture / solver: whatever it recorded stays
                    failed += 1
                    print("  (reference test %s raised %s: %s)" % (tid, type(e).__name__, str(e)[:80]))
    return ran, failed


def signature(expr):
    """Expression text + variable shapes and values: two tests differentiating the same thing count once."""
    h = hashlib.sha256(str(expr).encode())
    for v in sorted(expr.variables(), key=lambda v: v.id):
        h.update(str(v.shape).encode())
        h.update(np.ascontiguousarray(np.asarray(v.value, dtype=np.float64)).tobytes() if v.value is not None else b"none")
    h.update(str(expr.shape).encode())
    return h.hexdigest()


def main():
    orig_j, orig_h = _wrap("jacobian", "jac"), _wrap("hess_vec", "hess")
    ran, failed = run_reference_tests()
    Atom.jacobian, Atom.hess_vec = orig_j, orig_h
    print("reference rule tests executed: %d ok, %d raised; top-level calls recorded: %d" % (ran, failed, len(RECORDED)))
    seen, kept = {}, []
    for tid, kind, expr in RECORDED:
        if any(v.value is None for v in expr.variables()) or not expr.variables():
            continue
        sig = signature(expr)
        if sig in seen:
            seen[sig][1].add(kind)
            continue
        seen[sig] = (len(kept), {kind})
        kept.append((tid, expr))
    print("distinct expressions: %d" % len(kept))
    mga.OUT = OUT
    os.makedirs(OUT, exist_ok=True)
    for old in glob.glob(os.path.join(OUT, "*.npz")):
        os.remove(old)
    index = []
    for i, (tid, expr) in enumerate(kept):
        name = "ref_%03d" % i
        try:
            res = mga.make_one(name, lambda e=expr: (0, [e == 0]))
            index.append("%s\t%s\t%s\tjac=%s hess=%s %s" % ((name, tid, str(expr)[:100]) + res))
        except Exception as e:
            index.append("%s\t%s\t%s\tGENERATOR ERROR %s: %s" % (name, tid, str(expr)[:100], type(e).__name__, e))
        print(index[-1])
    with open(os.path.join(OUT, "INDEX.tsv"), "w") as f:
        f.write("fixture\treference test (first to differentiate it)\texpression\toutcome in the reference's Oracles\n")
        f.write("\n".join(index) + "\n")


if __name__ == "__main__":
    main()
