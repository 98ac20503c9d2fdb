"""Loader for the golden fixtures written by tests/golden/make_golden.py."""
import glob
import os

import numpy as np

from dnlp_b200 import ir

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


REFPROBLEMS_DIR = os.path.join(GOLDEN_DIR, "refproblems")


def refproblem_golden_names():
    """Fixtures of tests/golden/make_golden_refproblems.py: the problems of the reference's own NLP test-suite, taken
    from its test functions at their first ``solve(nlp=True)`` (INDEX.tsv there names the test behind each)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(REFPROBLEMS_DIR, "*.npz")))


class Golden:
    def __init__(self, name, directory=GOLDEN_DIR):
        z = np.load(os.path.join(directory, name + ".npz"), allow_pickle=False)
        self.name = name
        arrays = {k[3:]: z[k] for k in z.files if k.startswith("ir_a")}
        self.problem = ir.load_problem(str(z["ir_json"]), arrays)
        self.jac_rows, self.jac_cols = z["jac_rows"], z["jac_cols"]
        self.hess_rows, self.hess_cols = z["hess_rows"], z["hess_cols"]
        self.points = []
        for i in range(int(z["npoints"])):
            self.points.append({k: z["%s_%d" % (k, i)] for k in
                                ("x", "lam", "sigma", "f", "grad", "g", "jac", "hess")})


def assert_close(got, want, what, rtol=1e-10, atol=1e-12):
    """Oracle values must agree within rel 1e-10 in fp64 (BASELINE.json north_star)."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    ok = np.isclose(got, want, rtol=rtol, atol=atol, equal_nan=True)
    if not ok.all():
        bad = np.where(~ok)[0][:5]
        raise AssertionError("%s mismatch at %s: got %s want %s" % (what, bad, got[bad], want[bad]))


ATOMS_DIR = os.path.join(GOLDEN_DIR, "atoms")


def atom_golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(ATOMS_DIR, "*.npz")))


REFTESTS_DIR = os.path.join(GOLDEN_DIR, "reftests")


def reftest_golden_names():
    """Fixtures of tests/golden/make_golden_reftests.py: every expression the reference's own jacobian / hess_vec
    unit tests differentiate (INDEX.tsv there names the reference test behind each)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(REFTESTS_DIR, "*.npz")))


class AtomGolden:
    """Fixtures written by tests/golden/make_golden_atoms.py (raw rules, no Dnlp2Smooth)."""

    def __init__(self, name, directory=ATOMS_DIR):
        z = np.load(os.path.join(directory, name + ".npz"), allow_pickle=False)
        self.name = name
        self.jac_error, self.hess_error = str(z["jac_error"]), str(z["hess_error"])
        arrays = {k[3:]: z[k] for k in z.files if k.startswith("ir_a")}
        self.problem = ir.load_problem(str(z["ir_json"]), arrays)
        if not self.jac_error:
            self.jac_rows, self.jac_cols = z["jac_rows"], z["jac_cols"]
        if not self.hess_error:
            self.hess_rows, self.hess_cols = z["hess_rows"], z["hess_cols"]
        self.points = []
        for i in range(int(z["npoints"])):
            self.points.append({k: z["%s_%d" % (k, i)] for k in ("x", "lam", "sigma", "f", "grad", "g", "jac", "hess")
                                if "%s_%d" % (k, i) in z.files})
