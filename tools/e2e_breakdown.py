#!/usr/bin/env python
"""Where does an end-to-end evaluation (host buffers in/out) spend its time?  Per-callback wall
times of GpuOracles on one workload, next to the bytes each callback moves."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload, eval_point  # noqa: E402
from dnlp_b200.oracles import GpuOracles  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
prob = build_workload(name)
eager = None if len(sys.argv) < 3 else bool(int(sys.argv[2]))
o = GpuOracles(prob, eager=eager)
x, lam, sigma = eval_point(prob, 0)
rng = np.random.default_rng(7)
xs = [x * (1.0 + 1e-3 * rng.standard_normal(prob.n)) for _ in range(4)]
lams = [lam * (1.0 + 1e-3 * rng.standard_normal(prob.m)) for _ in range(4)]
calls = [("objective", lambda xi, li: o.objective(xi)), ("gradient", lambda xi, li: o.gradient(xi)),
         ("constraints", lambda xi, li: o.constraints(xi)), ("jacobian", lambda xi, li: o.jacobian(xi)),
         ("hessian", lambda xi, li: o.hessian(xi, li, sigma))]
for i in range(2):
    for _, fn in calls:
        fn(xs[i], lams[i])
tot = {k: 0.0 for k, _ in calls}
reps = 8
t_all = time.perf_counter()
for i in range(reps):
    for k, fn in calls:
        t0 = time.perf_counter()
        fn(xs[i % 4], lams[i % 4])
        tot[k] += time.perf_counter() - t0
t_all = time.perf_counter() - t_all
print("eager delivery: %s" % o.eager)
print("%s: n=%d m=%d nnzJ=%d nnzH=%d  cores=%d  dyn=%s" % (name, prob.n, prob.m, o.nnz_jac, o.nnz_hess, os.cpu_count(),
      {k: int(v[0].size) for k, v in o._dyn.items()}))
for k, _ in calls:
    print("  %-12s %8.3f ms" % (k, tot[k] / reps * 1e3))
print("  total        %8.3f ms  -> %.1f evals/s (fresh x and lambda every step)" % (t_all / reps * 1e3, reps / t_all))
# raw copies for reference
import ctypes as C
from dnlp_b200 import _cabi
a, h = _cabi.pinned_empty(prob.n)
t0 = time.perf_counter()
for i in range(5):
    a[:] = xs[i % 4]
print("  numpy copy of x into pinned: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for i in range(5):
    np.array_equal(a, xs[i % 4])
print("  numpy array_equal(x, x'):    %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
o.close()
