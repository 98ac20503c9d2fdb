#!/usr/bin/env python
"""Per-callback wall times of the row-sharded C3 oracle (host buffers in and out), every rank.
Launch with torchrun (one process per GPU):  python -m torch.distributed.run --nproc-per-node N ... this.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import Group, eval_point, sizes_of  # noqa: E402
from dnlp_b200 import workloads as W  # noqa: E402
from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_logistic_regression  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
grp = Group(rank, world, local_rank)
s = sizes_of("c3", float(os.environ.get("SCALE", "1.0")))
At, x_init = W.logistic_data(s["m"], s["n"], s["k"])
glob = W.logistic_regression(At, x_init)
local, layout = shard_logistic_regression(At, x_init, rank, world)
o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), store=grp.store, device=local_rank)
x, lam, sigma = eval_point(glob, 0)
rng = np.random.default_rng(7)
xs = [x * (1.0 + 1e-3 * rng.standard_normal(glob.n)) for _ in range(4)]
lams = [lam * (1.0 + 1e-3 * rng.standard_normal(glob.m)) for _ in range(4)]
calls = [("objective", lambda xi, li: o.objective(xi)), ("gradient", lambda xi, li: o.gradient(xi)),
         ("constraints", lambda xi, li: o.constraints(xi)), ("jacobian", lambda xi, li: o.jacobian(xi)),
         ("hessian", lambda xi, li: o.hessian(xi, li, sigma))]
for i in range(3):
    for _, fn in calls:
        fn(xs[i], lams[i])
grp.barrier()
tot = {k: 0.0 for k, _ in calls}
reps = 16
t_all = time.perf_counter()
for i in range(reps):
    for k, fn in calls:
        t0 = time.perf_counter()
        fn(xs[i % 4], lams[i % 4])
        tot[k] += time.perf_counter() - t0
t_all = time.perf_counter() - t_all
grp.barrier()
for r in range(world):
    if r == rank:
        print("rank %d of %d: shared-host outputs %s, local n=%d m=%d" % (rank, world, sorted(o._dev.shared), local.n, local.m))
        for k, _ in calls:
            print("  %-12s %8.3f ms" % (k, tot[k] / reps * 1e3))
        print("  total        %8.3f ms  -> %.1f evals/s" % (t_all / reps * 1e3, reps / t_all), flush=True)
    grp.barrier()
o.close()
grp.close()
