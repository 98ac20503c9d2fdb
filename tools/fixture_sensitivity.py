"""How close to the tolerance do the harvested reference-suite fixtures sit?  (Run on CPU; test infrastructure.)

The GPU computes the same tape with different rounding: fused multiply-adds, CUDA's own exp / log / pow (1-2 ulp),
tree-shaped and atomics-free but differently ordered row sums.  This tool replays every fixture of
tests/golden/refproblems and tests/golden/reftests through the NumPy tape interpreter with that kind of deviation
exaggerated - every elementwise result and every product term perturbed by up to ULPS ulp (default 16), the terms of
every row summed in a random order - and applies the tolerance of tests/test_zz_gpu_reference_suite.py.  A fixture that
fails here would be a marginal GPU test (a value that is the difference of large terms), not a kernel bug.

    python tools/fixture_sensitivity.py [ulps] [seed]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import tape_interp as TI  # noqa: E402
from dnlp_b200 import tape as T  # noqa: E402
from dnlp_b200.compiler import compile_problem  # noqa: E402
from golden_util import (REFPROBLEMS_DIR, REFTESTS_DIR, AtomGolden, Golden, assert_close,  # noqa: E402
                         refproblem_golden_names, reftest_golden_names)

ULPS = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
G_ATOL = 1e-9


def noise(a):
    a = np.asarray(a, dtype=float)
    return a * (1.0 + rng.integers(-ULPS, ULPS + 1, size=a.shape) * 2.0 ** -52)


_exact_f = TI._f
TI._f = lambda code, a, b, p: noise(_exact_f(code, a, b, p))


class NoisyInterp(TI.TapeInterp):
    def _run(self, prog, outs):
        t, V = self.t, self.V
        for i in prog:
            ins = t.instrs[i]
            if ins.kind != T.K_POLY:
                TI.TapeInterp._run(self, [i], outs)
                continue
            with np.errstate(all="ignore"):
                term = ins.coef.copy()
                m1 = ins.f1 >= 0
                term[m1] = term[m1] * V[ins.f1[m1]]
                m2 = ins.f2 >= 0
                term[m2] = term[m2] * V[ins.f2[m2]]
                term = noise(term)
                rows = np.repeat(np.arange(ins.count), np.diff(ins.ptr))
                perm = rng.permutation(term.size)
                res = np.zeros(ins.count)
                np.add.at(res, rows[perm], term[perm])
            if ins.dst_space == T.DST_V:
                V[ins.dst_off:ins.dst_off + ins.count] = res
            else:
                pos = np.arange(ins.count) if ins.pos is None else ins.pos
                if ins.accumulate:
                    outs[ins.dst_space][pos] += res
                else:
                    outs[ins.dst_space][pos] = res


def check(name, g, with_hessian=True):
    try:
        tape = compile_problem(g.problem, with_hessian=with_hessian)
    except Exception:
        return 0, 0                                       # rejected problems have no values to compare
    it, n, bad = NoisyInterp(tape), 0, 0
    for p in g.points:
        for k in ("f", "grad", "g", "jac") + (("hess",) if with_hessian else ()):
            got = it.eval(k, p["x"], p["lam"], float(p["sigma"])) if k == "hess" else it.eval(k, p["x"])
            n += 1
            try:
                assert_close(got, p[k], k, atol=G_ATOL if k == "g" else 1e-12)
            except AssertionError as e:
                bad += 1
                print("MARGINAL %s %s: %s" % (name, k, str(e)[:160]))
    return n, bad


if __name__ == "__main__":
    total = fails = 0
    for name in refproblem_golden_names():
        n, bad = check(name, Golden(name, REFPROBLEMS_DIR))
        total, fails = total + n, fails + bad
    for name in reftest_golden_names():
        g = AtomGolden(name, REFTESTS_DIR)
        if not g.jac_error:
            n, bad = check(name, g, with_hessian=not g.hess_error)
            total, fails = total + n, fails + bad
    print("ulps=%d: %d outputs checked, %d outside the tolerance" % (ULPS, total, fails))
    sys.exit(1 if fails else 0)
