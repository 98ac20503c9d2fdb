"""Digest of compiled tapes (every array and scalar of every instruction, patterns, constant parts, programs):
a refactoring of the compiler that is meant to be behaviour-preserving must leave these digests unchanged.

    python tools/tape_digest.py > before.txt;  <change the compiler>;  python tools/tape_digest.py | diff before.txt -

Covers every golden problem under tests/golden (incl. the harvested reference-suite problems) and scaled-down
instances of the five bench workloads (big enough to reach the large-problem emission paths)."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from dnlp_b200 import tape as T  # noqa: E402
from dnlp_b200.compiler import compile_problem  # noqa: E402


def _upd(h, v):
    if v is None:
        h.update(b"N")
    elif isinstance(v, np.ndarray):
        h.update(str((v.dtype.str, v.shape)).encode())
        h.update(np.ascontiguousarray(v).tobytes())
    elif isinstance(v, (list, tuple)):
        h.update(b"[")
        for x in v:
            _upd(h, x)
        h.update(b"]")
    elif isinstance(v, dict):
        for k in sorted(v):
            h.update(repr(k).encode())
            _upd(h, v[k])
    elif isinstance(v, (float, np.floating)):
        h.update(np.float64(v).tobytes())
    else:
        h.update(repr(v).encode())


def tape_digest(tape):
    h = hashlib.blake2b(digest_size=16)
    for ins in tape.instrs:
        for name in T.Instr.__slots__:
            h.update(name.encode())
            _upd(h, getattr(ins, name))
    for name in ("n", "m", "n_params", "nslots", "programs", "jac_rows", "jac_cols", "hess_rows", "hess_cols",
                 "jac_const", "hess_const", "grad_const", "g_const", "f_const", "jac_is_list", "dynamic",
                 "dynamic_sigma", "param_values"):
        h.update(name.encode())
        _upd(h, getattr(tape, name))
    return h.hexdigest()


def cases(scale):
    from golden_util import REFPROBLEMS_DIR, Golden, golden_names, refproblem_golden_names
    from dnlp_b200 import workloads as W
    for name in golden_names():
        yield "golden/" + name, (lambda name=name: Golden(name).problem)
    for name in refproblem_golden_names():
        yield "refproblems/" + name, (lambda name=name: Golden(name, REFPROBLEMS_DIR).problem)

    def c5():
        N = int(10_000_000 * scale) // 8 * 8
        return W.microbench(*W.microbench_data(N, N // 2, 10))

    def c3():
        return W.logistic_regression(*W.logistic_data(int(2_000_000 * scale), 4096, 16))
    yield "c2 n=%d" % int(8192 * scale ** 0.5), lambda: W.eigen_qcqp(int(8192 * scale ** 0.5))
    yield "c3 scale %g" % scale, c3
    yield "c5 scale %g" % scale, c5
    P, q, _ = W.qcqp_data(64, 4)
    yield "c4 n=64 k=4", lambda: W.qcqp(P, q)


if __name__ == "__main__":
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
    for label, make in cases(scale):
        prob = make()
        t = time.time()
        try:
            d = tape_digest(compile_problem(prob))
        except Exception as e:                                          # rejected problems must stay rejected
            d = "%s: %s" % (type(e).__name__, e)
        sys.stderr.write("%-50s %.2f s\n" % (label, time.time() - t))
        print(label, d)
