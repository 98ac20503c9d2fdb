#!/usr/bin/env python
"""Benchmark of the NLP oracle hot path: oracle evals/s for the set (f, grad f, g, J, Hess L).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c5|c1] [--impl reference]

One "step" = one evaluation of all five quantities at a fresh point.

  value     device-side throughput: (x, lambda, sigma) already resident in HBM, every per-x cache
            invalidated before each step, timed with CUDA events on the oracle's own stream.
  e2e       the same set through the public drop-in object (`GpuOracles`: the five cyipopt
            callbacks, pinned HOST buffers in and out, H2D/D2H inside the timed region).
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration, against the measured
            HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (oracle/dnlp_oracle.py: NumPy/SciPy restatement of the
            reference's algorithm) on a bounded sample of the same workload, on this box's host cores.

N > 1 (torchrun): the default workload does not shard a single evaluation; every rank evaluates its
own start point of the same problem (multi-start replicas, no collective) and the aggregate is
reported as weak scaling.  `--impl reference` times the CPU implementation (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROGS = ("f", "grad", "g", "jac", "hess")
# MEASURED_PEAKS.json has no fp64 figure; cuBLAS DGEMM 4096^3 on this pool's B200 (tools/cublas_dgemm_ref.py,
# profiles/r01_cublas_dgemm.txt) reached 35.4 TFLOP/s, 27.3 at the C4 shape [512x512]x[512x4096].
FP64_TENSOR_PEAK = 35.4
FP64_TENSOR_PEAK_SRC = "measured: cuBLAS DGEMM 4096^3 on this pool (profiles/r01_cublas_dgemm.txt); nominal 40"


# ------------------------------------------------------------------ workloads
def build_workload(name, scale=1.0):
    """Returns (ProblemIR, description dict).  `scale` < 1 shrinks the instance (CPU sample)."""
    from dnlp_b200 import workloads as W
    if name == "c1":
        return W.eigen_qcqp(3), {"workload": "c1: README toy eigen-QCQP n=3"}
    if name == "c2":
        n = max(8, int(round(8192 * scale)))
        return W.eigen_qcqp(n), {"workload": "c2: eigen-QCQP maximize quad_form(x,A) s.t. sum_squares(x)==1, "
                                 "dense A n=%d, single start" % n, "n": n}
    if name == "c3":
        m, n = max(64, int(2_000_000 * scale)), max(16, int(4096 * (scale ** 0.5)))
        At, x0 = W.logistic_data(m, n, 16)
        return W.logistic_regression(At, x0), {"workload": "c3: sparse logistic-type regression m=%d n=%d, "
                                               "16 nnz/row (lifted smooth form)" % (m, n), "m": m, "n": n}
    if name == "c5":
        N = max(64, int(10_000_000 * scale) // 8 * 8)
        m = max(8, N // 2)
        A, x0 = W.microbench_data(N, m, 10)
        return W.microbench(A, x0), {"workload": "c5: %d-node elementwise DAG + %d-nnz CSR constraint Jacobian"
                                     % (N, A.nnz), "N": N, "m": m, "nnz": int(A.nnz)}
    raise SystemExit("unknown workload %r" % name)


def describe_workload(name, scale=1.0):
    """The `config.workload` string of `build_workload(name, scale)` without generating the data."""
    if name == "c1":
        return "c1: README toy eigen-QCQP n=3"
    if name == "c2":
        return ("c2: eigen-QCQP maximize quad_form(x,A) s.t. sum_squares(x)==1, dense A n=%d, single start"
                % max(8, int(round(8192 * scale))))
    if name in ("c3", "c3s"):
        return ("c3: sparse logistic-type regression m=%d n=%d, 16 nnz/row (lifted smooth form)"
                % (max(64, int(2_000_000 * scale)), max(16, int(4096 * (scale ** 0.5)))))
    if name == "c4":
        return ("c4: multi-start batch of %d random starts of a nonconvex QCQP n=%d, 8 quadratic constraints; "
                "one eval = one start's full set" % (int(4096 * scale), max(16, int(512 * scale))))
    if name == "c5":
        N = max(64, int(10_000_000 * scale) // 8 * 8)
        return "c5: %d-node elementwise DAG + %d-nnz CSR constraint Jacobian" % (N, max(8, N // 2) * 10)
    return name


def eval_point(prob, rank, rng=None):
    rng = rng or np.random.default_rng(1000 + rank)
    x = np.asarray(prob.x0, dtype=np.float64) * (1.0 + 0.01 * rng.standard_normal(prob.n))
    lam = rng.standard_normal(prob.m)
    return x, lam, 1.0


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_port_evals_per_s(name, budget_s=20.0):
    """Time the CPU oracle port on a bounded sample of the workload; returns (evals/s scaled to the
    full workload, description)."""
    from oracle.dnlp_oracle import RefOracles
    if name == "c3s":
        name = "c3"
    if name == "c4":
        return cpu_port_c4(budget_s)
    sample_scale = {"c1": 1.0, "c2": 0.125, "c3": 0.01, "c5": 0.01}[name]
    prob, desc = build_workload(name, sample_scale)
    o = RefOracles(prob)
    o.jacobianstructure(), o.hessianstructure()
    x, lam, sigma = eval_point(prob, 0)
    reps, t_total = 0, 0.0
    with np.errstate(all="ignore"):
        while reps < 3 or (t_total < budget_s and reps < 50):
            t0 = time.perf_counter()
            o.objective(x), o.gradient(x), o.constraints(x), o.jacobian(x), o.hessian(x, lam, sigma)
            t_total += time.perf_counter() - t0
            reps += 1
            if t_total > budget_s:
                break
    per_eval = t_total / reps
    # cost is linear in the number of triplets; scale by the nnz ratio
    work_sample = o.jac_rows.size + 2 * o.hess_rows.size + prob.n + prob.m
    full = {"c1": 1.0, "c2": (8192 * 8193 // 2 * 2 + 3 * 8192) / max(work_sample, 1),
            "c3": 1.0 / 0.01, "c5": 1.0 / 0.01}[name]
    if name == "c2":
        n_s = desc["n"]
        full = (8192.0 / n_s) ** 2
    # the extrapolation assumes linear cost: check it on a second, half-size sample (SURVEY 8d asks
    # for the measured scaling next to extrapolated numbers)
    linearity = ""
    if name in ("c3", "c5"):
        prob2, _ = build_workload(name, sample_scale / 2)
        o2 = RefOracles(prob2)
        o2.jacobianstructure(), o2.hessianstructure()
        x2, lam2, _ = eval_point(prob2, 0)
        t2, r2 = 0.0, 0
        with np.errstate(all="ignore"):
            while r2 < 3 or (t2 < budget_s / 4 and r2 < 25):
                t0 = time.perf_counter()
                o2.objective(x2), o2.gradient(x2), o2.constraints(x2), o2.jacobian(x2), o2.hessian(x2, lam2, sigma)
                t2 += time.perf_counter() - t0
                r2 += 1
        linearity = "; half-size sample %.4f s/eval -> time ratio %.2f for a size ratio of 2" % (t2 / r2, per_eval / (t2 / r2))
    return 1.0 / (per_eval * full), {
        "sample": "%s; %d full evals in %.1f s (%.4f s/eval at sample size), scaled x%.1f by triplet count "
                  "to the full workload%s" % (desc["workload"], reps, t_total, per_eval, full, linearity),
        "seconds_per_eval_at_sample": per_eval, "scale_factor": full}


def cpu_port_c4(budget_s=10.0):
    from dnlp_b200 import workloads as W
    from oracle.dnlp_oracle import RefOracles
    P, q, rng = W.qcqp_data(512, 8)
    prob = W.qcqp(P, q)
    X = rng.uniform(-1, 1, (64, 512))
    lam = np.random.default_rng(17).standard_normal(8)
    r = RefOracles(prob)
    r.jacobianstructure(), r.hessianstructure()
    t0, cnt = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s or cnt < 3:
        xb = X[cnt % 64]
        r.objective(xb), r.gradient(xb), r.constraints(xb), r.jacobian(xb), r.hessian(xb, lam, 1.0)
        cnt += 1
    dt = time.perf_counter() - t0
    return cnt / dt, {"sample": "%d starts of the same QCQP (n=512, k=8) evaluated one by one in %.1f s" % (cnt, dt)}


# ------------------------------------------------------------------ row-sharded C3 (strong scaling)
def bench_sharded(args, rank, local_rank, world, dist, metric):
    """C3 with rows of A~ block-distributed over the ranks; one packed all-reduce per callback."""
    from dnlp_b200 import workloads as W
    from dnlp_b200.sharded import (GlobalStructure, RowShardedOracles, shard_logistic_regression,
                                   shard_microbench)
    if args.workload == "c3s":
        m, n = max(64, int(2_000_000 * args.scale)), max(16, int(4096 * (args.scale ** 0.5)))
        At, x_init = W.logistic_data(m, n, 16)
        glob = W.logistic_regression(At, x_init)
        local, layout = shard_logistic_regression(At, x_init, rank, world)
        wl = ("c3 row-sharded: sparse logistic-type regression m=%d n=%d, 16 nnz/row, rows of A block-distributed, "
              "lifted variables sharded, x replicated" % (m, n))
    else:
        N = max(64, int(10_000_000 * args.scale) // 8 * 8)
        m = max(8, N // 2)
        A5, x5 = W.microbench_data(N, m, 10)
        glob = W.microbench(A5, x5)
        local, layout = shard_microbench(A5, x5, rank, world)
        wl = ("c5 row-sharded: %d-node elementwise DAG + %d-nnz CSR Jacobian, constraint rows block-distributed, "
              "all variables replicated, Hessian contributions all-reduced (%d doubles)" % (N, A5.nnz, N))
    gs = GlobalStructure.from_problem(glob)
    comm = None
    if dist is None:
        class _Solo:
            rank, world = 0, 1

            def allreduce(self, v):
                return v
        comm = _Solo()
    o = RowShardedOracles(local, layout, gs, comm=comm, device=local_rank, root_only=True)
    rng = np.random.default_rng(3)
    x = glob.x0 * (1 + 0.01 * rng.standard_normal(glob.n))
    lam, sigma = rng.standard_normal(glob.m), 1.0

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
    # device side: every rank's local tape, max over ranks
    loc = o.local
    loc.upload_point(x[layout.var_map], lam[layout.con_map], sigma)
    loc.run_device(PROGS, args.warmup)
    barrier()
    ms = loc.run_device(PROGS, args.steps)
    barrier()
    xs = [x * (1 + 1e-6 * i) for i in range(4)]
    for i in range(2):
        o.objective(xs[i]), o.gradient(xs[i]), o.constraints(xs[i]), o.jacobian(xs[i]), o.hessian(xs[i], lam, sigma)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        xi = xs[i % 4]
        o.objective(xi), o.gradient(xi), o.constraints(xi), o.jacobian(xi), o.hessian(xi, lam, sigma)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        line = {"metric": metric, "value": args.steps / (ms * 1e-3), "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": wl, "parallelism": "row-sharded x%d, NCCL all-reduce + all-gather, device-side "
                           "assembly, outputs delivered to rank 0" % world,
                           "value_is": "max over ranks of the local tapes' CUDA-event time (collective excluded)"},
                "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "evals/s",
                        "api": "RowShardedOracles five callbacks, host buffers, all-reduce inside",
                        "h2d_bytes_per_step": int(5 * 8 * layout.var_map.size), "d2h_bytes_per_step": None},
                "gpu_launches": int(loc.kernel_launches())}
        print(json.dumps(line, default=float))
    o.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------ C4: batched multi-start QCQP
def bench_multistart(args, rank, local_rank, world, dist, metric, hbm_peak, peak_src):
    """4096 random starts of a nonconvex QCQP (n=512, 8 quadratic constraints); the batch is split
    evenly over the ranks (no collective).  One step = every start's (f, grad, g, J, Hess L)."""
    from dnlp_b200 import workloads as W
    from dnlp_b200.multistart import BatchedOracles
    n, k = max(16, int(512 * args.scale)), 8
    Btot = int(os.environ.get("DNLP_C4_BATCH", max(world, int(4096 * args.scale))))
    P, q, rng = W.qcqp_data(n, k)
    prob = W.qcqp(P, q)
    X = rng.uniform(-1, 1, (Btot, n))                       # the same rng stream as SURVEY 8(d) C4
    b0, b1 = (Btot * rank) // world, (Btot * (rank + 1)) // world
    B = b1 - b0
    lrng = np.random.default_rng(17)
    LAM = lrng.standard_normal((Btot, k))[b0:b1]
    SIG = np.ones(B)
    o = BatchedOracles(prob, B, device=local_rank)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
    o.upload(X[b0:b1], LAM, SIG)
    o.run_device(PROGS, args.warmup)
    l0 = o.kernel_launches()
    barrier()
    with ClockSampler(local_rank) as clk:
        ms = o.run_device(PROGS, args.steps)
        launches = o.kernel_launches() - l0
        barrier()
        o.eval(X[b0:b1], LAM, SIG)
        barrier()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            o.eval(X[b0:b1], LAM, SIG)
        e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        per = o.profile_instrs("all", iters=2)
        if os.environ.get("DNLP_BENCH_PROFILE"):
            for i in o.tape.programs["all"]:
                ii = o.tape.instrs[i]
                sys.stderr.write("[binstr %3d] kind=%d dst=%d rows=%-8d terms=%-9d %8.4f ms\n" % (
                    i, ii.kind, ii.dst_space, ii.count, 0 if ii.coef is None else ii.coef.size, per[i]))
        top = int(np.argmax(per))
        ins = o.tape.instrs[top]
        if ins.kind == 3:      # GEMM on the FP64 tensor cores
            flops = 2.0 * ins.count * ins.ncols * B
            roof = {"bound": "tensor", "kernel": "bgemm_dmma_kernel (instr %d: [%dx%d]x[%dx%d])" % (top, ins.count, ins.ncols, ins.ncols, B),
                    "achieved": flops / (per[top] * 1e-3) / 1e12, "peak": FP64_TENSOR_PEAK, "unit": "TFLOP/s",
                    "frac": flops / (per[top] * 1e-3) / 1e12 / FP64_TENSOR_PEAK, "traffic": None,
                    "peak_source": FP64_TENSOR_PEAK_SRC}
        else:
            nb = ins.nbytes_algorithmic() if ins.kind != 2 else (8 * ins.count * B * (2 if ins.accumulate else 1)
                                                                 + 12 * int(ins.coef.size))
            roof = {"bound": "hbm", "kernel": "batched instr %d kind %d (%d rows x %d starts)" % (top, ins.kind, ins.count, B),
                    "achieved": nb / (per[top] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": nb / (per[top] * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes": int(nb)}
        roof["ms"] = float(per[top])
        roof["share_of_step"] = float(per[top] / max(per.sum(), 1e-12))
        gemm_ms = float(sum(per[i] for i in o.tape.programs["all"] if o.tape.instrs[i].kind == 3))
        gemm_fl = float(sum(2.0 * o.tape.instrs[i].count * o.tape.instrs[i].ncols * B
                            for i in o.tape.programs["all"] if o.tape.instrs[i].kind == 3))
        line = {"metric": metric, "value": Btot * args.steps / (ms * 1e-3), "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "c4: multi-start batch of %d random starts of a nonconvex QCQP n=%d, %d quadratic "
                           "constraints; one eval = one start's full set" % (Btot, n, k),
                           "parallelism": "starts split over %d GPU(s), no collective" % world,
                           "l2": "outputs larger than L2", "starts_per_gpu": B},
                "dmma": {"gemm_tflops_individual_launches": gemm_fl / max(gemm_ms * 1e-3, 1e-12) / 1e12,
                         "gemm_ms_individual_launches": gemm_ms,
                         "gemm_tflops_grouped_launch_est": gemm_fl / max((ms / args.steps - (float(per.sum()) - gemm_ms)) * 1e-3,
                                                                         1e-12) / 1e12,
                         "note": "the k+1 independent maps run as ONE grouped grid inside the step; its time is "
                                 "estimated as step time minus the other instructions' times",
                         "ncu_dmma_pipe_pct_individual": 75.1,
                         "peak_measured_tflops": FP64_TENSOR_PEAK, "peak_source": FP64_TENSOR_PEAK_SRC,
                         "cublas_same_shape_tflops": 27.3},
                "clocks": clk.summary(),
                "e2e": {"value": Btot * e_steps / (e2e_ms * 1e-3), "unit": "evals/s",
                        "api": "BatchedOracles.eval (host arrays in/out)",
                        "h2d_bytes_per_step": int(8 * B * (n + k + 1)),
                        "d2h_bytes_per_step": int(8 * B * (1 + n + k + o.nnz_jac + o.nnz_hess))},
                "gpu_launches": int(launches), "roofline": roof}
        if not args.no_cpu_baseline:
            from oracle.dnlp_oracle import RefOracles
            r = RefOracles(prob)
            r.jacobianstructure(), r.hessianstructure()
            t0, cnt = time.perf_counter(), 0
            while time.perf_counter() - t0 < 10.0:
                xb = X[cnt % Btot]
                r.objective(xb), r.gradient(xb), r.constraints(xb), r.jacobian(xb), r.hessian(xb, LAM[0], 1.0)
                cnt += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": cnt / dt, "unit": "evals/s", "cores": 1, "kind": "port",
                                    "sample": "%d starts of the same QCQP evaluated one by one in %.1f s" % (cnt, dt)}
        print(json.dumps(line, default=float))
    o.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("DNLP_BENCH_WORKLOAD", "c2"))
    ap.add_argument("--scale", type=float, default=float(os.environ.get("DNLP_BENCH_SCALE", "1.0")))
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="profiling aid: only the device-resident step loop and the per-instruction timing "
                         "(the ncu launch list then holds the step's kernels and nothing else)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    metric = "oracle evals/s (f, grad f, g, J, Hess L)"

    if args.impl == "reference":
        if rank != 0:
            return
        v, d = cpu_port_evals_per_s(args.workload, budget_s=max(5.0, min(60.0, 2.0 * args.steps)))
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": describe_workload(args.workload, args.scale),
                           "parallelism": "CPU oracle port, single process"},
                "cpu_baseline": {"value": v, "unit": "evals/s", "cores": 1, "kind": "port", "sample": d["sample"]},
                "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    if args.workload in ("c3s", "c5s"):
        return bench_sharded(args, rank, local_rank, world, dist, metric)
    if args.workload == "c4":
        return bench_multistart(args, rank, local_rank, world, dist, metric, hbm_peak, peak_src)

    from dnlp_b200.oracles import GpuOracles
    prob, desc = build_workload(args.workload, args.scale)
    t0 = time.time()
    tape, tape_file = None, None
    if os.environ.get("DNLP_TAPE_CACHE"):
        # profiling aid: repeated invocations on one box (bench, ncu launch list, ncu full capture) reuse
        # the compiled tape instead of paying the DAG compiler again (55 s for c5)
        import pickle
        tape_file = os.path.join(os.environ["DNLP_TAPE_CACHE"], "tape_%s_%g.pkl" % (args.workload, args.scale))
        if os.path.exists(tape_file):
            tape = pickle.load(open(tape_file, "rb"))
    o = GpuOracles(prob, device=local_rank, tape=tape)
    compile_s = time.time() - t0
    if tape_file and tape is None:
        import pickle
        pickle.dump(o.tape, open(tape_file, "wb"), protocol=4)
    if rank == 0:
        import resource
        sys.stderr.write("[bench] %s: compiled in %.1f s, n=%d m=%d nnzJ=%d nnzH=%d, %d instructions, "
                         "host maxrss %.1f GB\n"
                         % (desc["workload"], compile_s, prob.n, prob.m, o.nnz_jac, o.nnz_hess, len(o.tape.instrs),
                            resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6))
    x, lam, sigma = eval_point(prob, rank)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    o.upload_point(x, lam, sigma)
    o.run_device(PROGS, args.warmup)
    launches0 = o.kernel_launches()
    barrier()
    with ClockSampler(local_rank) as clk:
        ms = o.run_device(PROGS, args.steps)
        launches = o.kernel_launches() - launches0
        if rank == 0:
            sys.stderr.write("[bench] device: %.4f ms/eval\n" % (ms / args.steps))
        barrier()
        e2e_s = fused_s = sv_s = float("nan")
        e2e_steps = args.steps
        cb_dev, cb_e2e = {}, {}
        if not args.device_only:
            # ---- end to end through the public callbacks, host buffers --------------------------------
            rng = np.random.default_rng(7 + rank)
            npts = min(args.steps, 4)
            xs = [x * (1.0 + 1e-3 * rng.standard_normal(prob.n)) for _ in range(npts)]
            lams = [lam * (1.0 + 1e-3 * rng.standard_normal(prob.m)) for _ in range(npts)]

            def five(i, sg):
                xi, li = xs[i % npts], lams[i % npts]
                o.objective(xi), o.gradient(xi), o.constraints(xi), o.jacobian(xi), o.hessian(xi, li, sg)
            for i in range(2):
                five(i, sigma)
            barrier()
            e2e_steps = args.steps
            # headline e2e: a new x and a new lambda every step, the objective factor held at 1.0 - the way
            # IPOPT calls eval_h in every regular iteration (it passes 0 only in the restoration phase)
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                five(i, sigma)
            e2e_s = time.perf_counter() - t0
            barrier()
            # worst case for the sigma-keyed Hessian entries: the objective factor changes every step too
            sv_steps = max(2, min(e2e_steps, 6))
            five(0, 0.5)
            t0 = time.perf_counter()
            for i in range(sv_steps):
                five(i, 1.0 if i % 2 else 0.5)
            sv_s = (time.perf_counter() - t0) * e2e_steps / sv_steps
            barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                o.eval_all(xs[i % npts], lams[i % npts], sigma)
            fused_s = time.perf_counter() - t0
            # each callback on its own: device-side with nothing cached, and through the host API
            cb_iters = max(3, min(args.steps, 10))
            o.upload_point(x, lam, sigma)
            cb_dev = {p: o.run_device((p,), cb_iters) / cb_iters for p in PROGS}
            cb_e2e = {p: 0.0 for p in PROGS}
            fns = {"f": lambda xi, li: o.objective(xi), "grad": lambda xi, li: o.gradient(xi),
                   "g": lambda xi, li: o.constraints(xi), "jac": lambda xi, li: o.jacobian(xi),
                   "hess": lambda xi, li: o.hessian(xi, li, sigma)}
            for i in range(cb_iters):
                for p in PROGS:
                    t0 = time.perf_counter()
                    fns[p](xs[i % npts], lams[i % npts])
                    cb_e2e[p] += (time.perf_counter() - t0) * 1e3 / cb_iters
    clocks = clk.summary()

    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_s * 1e3, fused_s * 1e3, sv_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, fused_ms, sv_ms = [float(v) for v in t.tolist()]
    else:
        e2e_ms, fused_ms, sv_ms = e2e_s * 1e3, fused_s * 1e3, sv_s * 1e3

    if rank == 0:
        # ---- roofline of the dominant kernel (CUDA events around every instruction) --------------
        per = o.profile_instrs("all", iters=3)
        if os.environ.get("DNLP_BENCH_PROFILE"):
            names = {1: "elem", 2: "poly", 3: "gemv", 4: "scale"}
            for i in o.tape.programs["all"]:
                ii = o.tape.instrs[i]
                nb = ii.nbytes_algorithmic()
                sys.stderr.write("[instr %3d] %-5s dst=%d rows=%-9d terms=%-9d %8.4f ms %8.1f GB/s\n" % (
                    i, names[ii.kind], ii.dst_space, ii.count, 0 if ii.coef is None else ii.coef.size,
                    per[i], nb / max(per[i], 1e-9) / 1e6))
        top = int(np.argmax(per))
        ins = o.tape.instrs[top]
        kind = {1: "elem", 2: "poly", 3: "gemv", 4: "scale"}[ins.kind]
        alg_bytes = ins.nbytes_algorithmic()
        achieved = alg_bytes / (per[top] * 1e-3) / 1e9 if per[top] > 0 else 0.0
        total_alg = sum(o.tape.instrs[i].nbytes_algorithmic() for i in o.tape.programs["all"])
        h2d = prob.n * 8 + (prob.m + 1) * 8        # x once per step (unchanged-point detection), lambda + sigma
        d2h_full = 8 * (1 + prob.n + prob.m + o.nnz_jac + o.nnz_hess)
        kname = o.instr_kernel(top)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
            traffic = tj.get(args.workload, {}).get(kname, {}).get("dram_bytes_per_launch_max")
        except Exception:
            pass
        full = {"grad": prob.n, "g": prob.m, "jac": o.nnz_jac, "hess": o.nnz_hess}
        d2h = 8 * (1 + sum(o._dyn[k][0].size if k in o._dyn else v for k, v in full.items()))
        line = {
            "metric": metric, "value": world * args.steps / (ms * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(desc, parallelism="single start" if world == 1 else
                           "multi-start replicas: one start point per GPU, no collective",
                           l2="inputs larger than L2 (126 MB)" if total_alg > 2 * 126e6 else
                           "working set fits L2; no flush", nnz_jac=o.nnz_jac, nnz_hess=o.nnz_hess,
                           compile_s=round(compile_s, 2), tape_from_cache=tape is not None,
                           algorithmic_bytes_per_eval=int(total_alg)),
            "hbm_gbs_whole_eval": total_alg / (ms / args.steps * 1e-3) / 1e9,
            "clocks": clocks,
            "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "GpuOracles.objective/gradient/constraints/jacobian/hessian (5 callbacks)",
                    "inputs": "new x and new lambda every step, sigma = 1.0 (IPOPT's calling pattern)",
                    "elided": "entries that are compile-time constants (affine Jacobian rows) or depend on sigma only "
                              "(2*sigma*Q of a quad_form objective) stay in the reused host arrays",
                    "sigma_changing_every_step_value": world * e2e_steps / (sv_ms * 1e-3),
                    "d2h_bytes_per_step_without_elision": int(d2h_full),
                    "fused_eval_all_value": world * e2e_steps / (fused_ms * 1e-3)},
            "gpu_launches": int(launches),
            "per_callback": {"device_ms_nothing_cached": cb_dev, "e2e_ms_in_ipopt_call_order": cb_e2e,
                             "note": "device: each program alone at a fresh point (shared forward work is repeated); "
                                     "e2e: objective first, so it carries the upload of x and the shared sweep"},
            "roofline": {"bound": "hbm", "kernel": "%s (instr %d, %d rows)" % (kname or kind, top, ins.count),
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": "profiles/r01_ncu_traffic.json (ncu --set full, dram read+write)"
                         if traffic else None,
                         "algorithmic_bytes": int(alg_bytes), "ms": float(per[top]),
                         "peak_source": peak_src,
                         "share_of_step": float(per[top] / max(per.sum(), 1e-12))},
        }
        if args.device_only:
            line["e2e"] = None
            line["note"] = "--device-only profiling run: e2e legs skipped, not a bench line"
        if not args.no_cpu_baseline and not args.device_only:
            v, d = cpu_port_evals_per_s(args.workload)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": 1, "kind": "port",
                                    "sample": d["sample"]}
        print(json.dumps(line, default=float))
    o.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
