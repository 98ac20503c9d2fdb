"""Which instruction variants (= kernel template instances and launch branches of csrc/dnlp_cabi.cu) do the fixtures
harvested from the reference's suite need that the GPU-validated hand-written fixtures do not already launch?
Compile-only, runs on CPU.  Round 2 result: 12 of 106 signatures are new, every one a single-row POLY instruction
(poly_rows_kernel, count == 1) that differs from a validated one only in its destination array / pos / f2 flags.

    python tools/kernel_variant_coverage.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from dnlp_b200.compiler import compile_problem  # noqa: E402
from golden_util import (REFPROBLEMS_DIR, REFTESTS_DIR, AtomGolden, Golden, atom_golden_names, golden_names,  # noqa: E402
                         refproblem_golden_names, reftest_golden_names)

FIELDS = ("kind", "fcode", "has_f2", "has_unit_factor", "pos", "accumulate", "no_ptr", "one_term_rows",
          "empty_rows", "single_row", "dst_space", "dst_stride", "a_stride", "b_stride", "post_scale", "panel")


def sig(ins):
    has_f2 = ins.f2 is not None and bool(np.any(ins.f2 >= 0))
    unit = ins.f1 is not None and bool(np.any(ins.f1 < 0))
    lens = None if ins.ptr is None else np.diff(ins.ptr)
    return (ins.kind, ins.fcode, has_f2, unit, ins.pos is not None, bool(ins.accumulate), ins.ptr is None,
            lens is not None and ins.coef is not None and ins.coef.size == ins.count,
            lens is not None and bool(np.any(lens == 0)), ins.count == 1, ins.dst_space, ins.dst_stride, ins.a_stride,
            ins.b_stride, ins.post_scale != 1.0, ins.panel_prev is not None)


def sigs(problem, with_hessian=True):
    try:
        return {sig(i) for i in compile_problem(problem, with_hessian=with_hessian).instrs}
    except Exception:
        return set()


def atom_sigs(g):
    return set() if g.jac_error else sigs(g.problem, with_hessian=not g.hess_error)


if __name__ == "__main__":
    validated, harvested, where = set(), set(), {}
    for n in golden_names():
        validated |= sigs(Golden(n).problem)
    for n in atom_golden_names():
        validated |= atom_sigs(AtomGolden(n))
    for n in refproblem_golden_names():
        s = sigs(Golden(n, REFPROBLEMS_DIR).problem)
        harvested |= s
        for x in s:
            where.setdefault(x, []).append(n)
    for n in reftest_golden_names():
        s = atom_sigs(AtomGolden(n, REFTESTS_DIR))
        harvested |= s
        for x in s:
            where.setdefault(x, []).append(n)
    new = sorted(harvested - validated)
    print("validated fixtures: %d signatures; harvested: %d; new: %d" % (len(validated), len(harvested), len(new)))
    for x in new:
        print(dict(zip(FIELDS, x)), where[x][:4])
