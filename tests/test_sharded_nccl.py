"""Row-sharded oracle with the device-side exchange of csrc/dnlp_shard.cu (one process per rank).

* 2 and 3 ranks SHARING one GPU: the peer-memory path alone (CUDA IPC exchange areas, one-shot
  all-reduce kernels, direct stores into the root's global array) and the shared-host delivery of
  outputs that have nothing to sum (one host array for all ranks, every rank copies its own runs,
  every rank returns the full output) - runs on the 1-GPU test box;
* 2 ranks on 2 GPUs with NCCL as well (the ncclAllReduce route forced for the shared entries) -
  skipped when the box has one GPU.

Every callback of the root must equal the single-GPU oracle of the GLOBAL problem (structures bit for
bit, values rel 1e-10); the objective must be known on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, same_gpu, kind, allreduce, host_share=True):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["DNLP_SHARD_ALLREDUCE"] = allreduce
        os.environ["DNLP_SHARD_HOST_SHARE"] = "1" if host_share else "0"
        if host_share == "broken":                  # shm_open rejects the name: the setup must fall back, not fail
            os.environ["DNLP_SHARD_SHM_PREFIX"] = "/no/such/dir/dnlp"
            host_share = False
        from dnlp_b200 import workloads as W
        from dnlp_b200.comm import SocketStore, barrier
        from dnlp_b200.oracles import GpuOracles
        from dnlp_b200.sharded import (GlobalStructure, RowShardedOracles, shard_logistic_regression,
                                       shard_microbench)
        from golden_util import assert_close
        dev = 0 if same_gpu else rank
        store = SocketStore(rank, world, "127.0.0.1", port)
        if kind == "c3":
            At, x_init = W.logistic_data(20011, 48, 8, seed=5)
            glob = W.logistic_regression(At, x_init)
            local, layout = shard_logistic_regression(At, x_init, rank, world)
        else:                       # C5-type: every variable replicated, Hessian contributions summed
            A, x0 = W.microbench_data(4000, 1531, 6, seed=3)
            glob = W.microbench(A, x0)
            local, layout = shard_microbench(A, x0, rank, world)
        ref = GpuOracles(glob, device=dev)
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), store=store, device=dev,
                              nccl=not same_gpu)
        assert o._dev is not None, "device-side exchange not active"
        if allreduce == "nccl":
            assert o._dev.comm.has_nccl
        shared = set(o._dev.shared)
        if kind == "c3":       # gradient, constraints and Hessian of C3 have nothing to sum and contiguous owned runs
            assert shared == ({"grad", "g", "hess"} if host_share else set()), shared
        full = lambda name: rank == 0 or name in shared      # noqa: E731  (shared-host outputs are complete on EVERY rank)
        np.testing.assert_array_equal(o.jacobianstructure()[0], ref.jacobianstructure()[0])
        np.testing.assert_array_equal(o.jacobianstructure()[1], ref.jacobianstructure()[1])
        np.testing.assert_array_equal(o.hessianstructure()[0], ref.hessianstructure()[0])
        np.testing.assert_array_equal(o.hessianstructure()[1], ref.hessianstructure()[1])
        rng = np.random.default_rng(11)           # same stream on every rank: same global point
        for it in range(3):
            x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
            lam = rng.standard_normal(glob.m)
            sigma = float(rng.uniform(0.5, 1.5))
            f = o.objective(x)
            assert_close(f, ref.objective(x), "f")                   # every rank knows the objective
            got = [o.gradient(x), o.constraints(x), o.jacobian(x), o.hessian(x, lam, sigma)]
            if it == 1:                                              # the same output twice in a row (line search)
                got[1] = o.constraints(x)
            if full("grad"):
                assert_close(got[0], ref.gradient(x), "grad")
            if full("g"):
                assert_close(got[1], ref.constraints(x), "g", atol=1e-11)
            if full("jac"):
                assert_close(got[2], ref.jacobian(x), "jac")
            if full("hess"):
                assert_close(got[3], ref.hessian(x, lam, sigma), "hess")
        ms = o.run_device(iters=5)                        # reduce of the shared entries trails one evaluation behind
        assert ms > 0
        x = glob.x0 * 1.02                                # ... and the callbacks still agree afterwards
        assert_close(o.objective(x), ref.objective(x), "f after the device loop")
        hh = o.hessian(x, lam, 0.7)
        if full("hess"):
            assert_close(hh, ref.hessian(x, lam, 0.7), "hess after the device loop")
        barrier(store)
        o.close(), ref.close()
        if full("hess"):
            assert_close(o._out["hess"], hh, "the output array outlives the shard handle")
        store.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def _worker_loop(rank, world, port, q, same_gpu, kind, allreduce, host_share=True):
    """Worker-loop mode: ONLY rank 0 issues callbacks (as the one solver process would); the others serve."""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["DNLP_SHARD_HOST_SHARE"] = "1" if host_share else "0"
        from dnlp_b200 import workloads as W
        from dnlp_b200.comm import SocketStore
        from dnlp_b200.oracles import GpuOracles
        from dnlp_b200.sharded import (GlobalStructure, RowShardedOracles, shard_logistic_regression,
                                       shard_microbench)
        from golden_util import assert_close
        dev = 0 if same_gpu else rank
        store = SocketStore(rank, world, "127.0.0.1", port)
        if kind == "c3":
            At, x_init = W.logistic_data(20011, 48, 8, seed=5)
            glob = W.logistic_regression(At, x_init)
            local, layout = shard_logistic_regression(At, x_init, rank, world)
        else:
            A, x0 = W.microbench_data(4000, 1531, 6, seed=3)
            glob = W.microbench(A, x0)
            local, layout = shard_microbench(A, x0, rank, world)
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), store=store, device=dev,
                              nccl=not same_gpu, workers=True)
        ncalls = 0
        if rank == 0:
            ref = GpuOracles(glob, device=dev)
            rng = np.random.default_rng(23)
            for it in range(4):
                x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
                lam = rng.standard_normal(glob.m)
                sigma = float(rng.uniform(0.5, 1.5))
                # IPOPT's order at an iterate, a line-search style repeat, a Knitro-style list for x
                assert_close(o.objective(x), ref.objective(x), "f")
                assert_close(o.constraints(list(x) if it == 2 else x), ref.constraints(x), "g", atol=1e-11)
                assert_close(o.gradient(x), ref.gradient(x), "grad")
                assert_close(o.jacobian(x), ref.jacobian(x), "jac")
                assert_close(o.hessian(x, lam, sigma), ref.hessian(x, lam, sigma), "hess")
                assert_close(o.hessian(x, 2 * lam, 0.0), ref.hessian(x, 2 * lam, 0.0), "hess, new multipliers")
                ncalls += 6
            with pytest.raises(RuntimeError):
                o.serve()                                    # the root runs the solver
            o.release_workers()
        else:
            served = o.serve()
            assert served == 24, served
        # every rank calls for a while (SPMD) at ANOTHER point, then the loop resumes at the point posted last: the
        # shared copy still holds it, the devices do not - the root must re-stage, not trust the compare
        from dnlp_b200.comm import barrier
        rng2 = np.random.default_rng(99)
        x2 = glob.x0 * (1 + 0.05 * rng2.standard_normal(glob.n))
        lam2 = rng2.standard_normal(glob.m)
        o.set_worker_loop(False)
        barrier(store)
        f2, h2 = o.objective(x2), o.hessian(x2, lam2, 1.25)
        o.set_worker_loop(True)
        if rank == 0:
            assert_close(f2, ref.objective(x2), "f (SPMD phase)")
            assert_close(h2, ref.hessian(x2, lam2, 1.25), "hess (SPMD phase)")
            assert_close(o.objective(x), ref.objective(x), "f after the switch back")
            assert_close(o.hessian(x, 2 * lam, 0.0), ref.hessian(x, 2 * lam, 0.0), "hess after the switch back")
            assert_close(o.gradient(x), ref.gradient(x), "grad after the switch back")
            ref.close()
            o.close()                                        # ... and releases the workers
        else:
            assert o.serve() == 3
            o.close()
        store.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def _run(world, same_gpu, kind, allreduce="auto", host_share=True, target=None):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target or _worker, args=(r, world, port, q, same_gpu, kind, allreduce, host_share))
             for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    results, deadline, problem = [], time.time() + 300, None
    while len(results) < world and problem is None:
        try:
            results.append(q.get(timeout=2))
            if results[-1][1] != "ok":                 # the peers of a failed rank would wait for it: stop them
                problem = "rank %d: %s" % results[-1]
        except queue.Empty:
            done = {r for r, _ in results}
            dead = [(i, p.exitcode) for i, p in enumerate(procs) if p.exitcode not in (None, 0) and i not in done]
            if dead:
                problem = "rank(s) died without reporting (rank, exit code): %s" % dead
            elif time.time() > deadline:
                problem = "no report within 300 s"
    for p in procs:
        if problem is not None and p.is_alive():
            p.terminate()
        p.join(timeout=60)
    assert problem is None, problem
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


@pytest.mark.parametrize("world,kind", [(2, "c3"), (3, "c3"), (2, "c5")])
def test_row_sharded_peer_memory_exchange_on_one_gpu(world, kind):
    _run(world, True, kind)


def test_row_sharded_nvlink_delivery_without_the_shared_host_array():
    """DNLP_SHARD_HOST_SHARE=0: owners store into the root's device array, one D2H leaves the root."""
    _run(2, True, "c3", host_share=False)


def test_row_sharded_falls_back_when_the_shared_segment_cannot_be_created():
    _run(2, True, "c3", host_share="broken")


@pytest.mark.parametrize("world,kind,host_share", [(2, "c3", True), (3, "c3", True), (2, "c5", True), (2, "c3", False)])
def test_worker_loop_only_the_root_issues_callbacks(world, kind, host_share):
    """One solver process (ipopt_nlpif.py:143-170 drives the callbacks from one cyipopt.Problem): rank 0 calls the
    seven callbacks, the other ranks sit in serve() and follow through the shared host segments."""
    _run(world, True, kind, host_share=host_share, target=_worker_loop)


@pytest.mark.parametrize("kind,allreduce", [("c3", "auto"), ("c5", "auto"), ("c5", "nccl")])
def test_row_sharded_two_gpus(kind, allreduce):
    from dnlp_b200 import _cabi
    if _cabi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, False, kind, allreduce)
