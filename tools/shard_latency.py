#!/usr/bin/env python
"""Why does one evaluation of a 1/W row shard of C3 take what it takes?  Runs the LOCAL tape of one
rank of a W-way split on a single GPU (no exchange: the shared objective is one double) and prints the
whole-evaluation time with graphs / parallel lanes on and off next to the per-instruction times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import sizes_of  # noqa: E402
from dnlp_b200 import workloads as W  # noqa: E402
from dnlp_b200.oracles import GpuOracles  # noqa: E402
from dnlp_b200.sharded import shard_logistic_regression  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = sizes_of("c3", 1.0)
At, x_init = W.logistic_data(s["m"], s["n"], s["k"])
local, layout = shard_logistic_regression(At, x_init, rank, world)
o = GpuOracles(local)
rng = np.random.default_rng(0)
o.upload_point(local.x0, rng.standard_normal(local.m), 1.0)
PROGS = ("f", "grad", "g", "jac", "hess")
iters = 2000
print("C3 shard %d of %d: n=%d m=%d instrs in the union program: %d" % (rank, world, local.n, local.m,
                                                                        len(o.tape.programs["all"])))
for graphs in (True, False):
    for par in (True, False):
        o.set_graphs(graphs), o.set_parallel(par)
        o.run_device(PROGS, 50)
        ms = o.run_device(PROGS, iters) / iters
        print("  graphs=%d parallel=%d: %.4f ms per evaluation" % (graphs, par, ms))
o.set_graphs(True), o.set_parallel(True)
per = o.profile_instrs("all", iters=20)
names = {1: "elem", 2: "poly", 3: "gemv", 4: "scale", 5: "spmvj"}
tot = 0.0
for i in o.tape.programs["all"]:
    ii = o.tape.instrs[i]
    tot += per[i]
    print("  instr %3d %-5s dst=%d rows=%-8d terms=%-9d %7.4f ms  %s" % (
        i, names.get(ii.kind, "?"), ii.dst_space, ii.count, 0 if ii.coef is None else ii.coef.size, per[i],
        o.instr_kernel(i)))
print("  sum of instruction times (each alone, CUDA events): %.4f ms" % tot)
for p in PROGS:
    o.run_device((p,), 50)
    print("  program %-5s alone: %.4f ms" % (p, o.run_device((p,), iters) / iters))
o.close()
