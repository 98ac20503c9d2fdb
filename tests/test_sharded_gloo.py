"""Row-sharded oracle, world_size 2 and 3 on CPU: over gloo (torch.distributed, through the store
adapter below) and over the package's own stdlib rendezvous (dnlp_b200.comm.SocketStore).

Checks the host logic of dnlp_b200.sharded (index maps into the global reference pattern,
dynamic-entry compaction, sum over ranks).  The local evaluator is the CPU oracle here (test
infrastructure); on GPUs it is GpuOracles with the device-side exchange (tests/test_sharded_nccl.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class GlooStore:
    """The store interface of dnlp_b200.comm on top of a torch.distributed (gloo) group."""

    def __init__(self, dist):
        self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()

    def allgather(self, payload):
        out = [None] * self.world
        self.dist.all_gather_object(out, bytes(payload))
        return out


def _make_store(kind, rank, world, port):
    if kind == "gloo":
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        return GlooStore(dist), dist
    from dnlp_b200.comm import SocketStore
    return SocketStore(rank, world, "127.0.0.1", port), None


def _worker_c5(rank, world, port, q, kind="gloo"):
    """C5-type sharding: every variable replicated, Hessian contributions of all ranks summed."""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        store, dist = _make_store(kind, rank, world, port)
        from dnlp_b200 import workloads as W
        from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_microbench
        from golden_util import assert_close
        from oracle.dnlp_oracle import RefOracles

        A, x0 = W.microbench_data(160, 53, 6, seed=3)
        glob = W.microbench(A, x0)
        ref = RefOracles(glob)
        jr, jc = ref.jacobianstructure()
        hr, hc = ref.hessianstructure()
        local, layout = shard_microbench(A, x0, rank, world)
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), store=store, oracle_factory=RefOracles)
        np.testing.assert_array_equal(o.jacobianstructure()[0], jr)
        np.testing.assert_array_equal(o.jacobianstructure()[1], jc)
        np.testing.assert_array_equal(o.hessianstructure()[0], hr)
        rng = np.random.default_rng(4)
        for _ in range(2):
            x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
            lam = rng.standard_normal(glob.m)
            assert_close(o.objective(x), ref.objective(x), "f")
            assert_close(o.gradient(x), ref.gradient(x), "grad")
            assert_close(o.constraints(x), ref.constraints(x), "g")
            assert_close(o.jacobian(x), ref.jacobian(x), "jac")
            assert_close(o.hessian(x, lam, 0.8), ref.hessian(x, lam, 0.8), "hess")
        store.allgather(b"")
        if dist is not None:
            dist.destroy_process_group()
        else:
            store.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def _worker(rank, world, port, q, kind="gloo"):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        store, dist = _make_store(kind, rank, world, port)
        from dnlp_b200 import workloads as W
        from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_logistic_regression
        from golden_util import assert_close
        from oracle.dnlp_oracle import RefOracles

        At, x_init = W.logistic_data(301, 12, 4, seed=5)
        glob = W.logistic_regression(At, x_init)
        ref = RefOracles(glob)
        jr, jc = ref.jacobianstructure()
        hr, hc = ref.hessianstructure()
        local, layout = shard_logistic_regression(At, x_init, rank, world)
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), store=store, oracle_factory=RefOracles)
        np.testing.assert_array_equal(o.jacobianstructure()[0], jr)
        np.testing.assert_array_equal(o.jacobianstructure()[1], jc)
        np.testing.assert_array_equal(o.hessianstructure()[0], hr)
        np.testing.assert_array_equal(o.hessianstructure()[1], hc)
        rng = np.random.default_rng(11)           # same stream on every rank: same global x
        for _ in range(3):
            x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
            lam = rng.standard_normal(glob.m)
            sigma = float(rng.uniform(0.5, 1.5))
            assert_close(o.objective(x), ref.objective(x), "f")
            assert_close(o.gradient(x), ref.gradient(x), "grad")
            assert_close(o.constraints(x), ref.constraints(x), "g")
            assert_close(o.jacobian(x), ref.jacobian(x), "jac")
            assert_close(o.hessian(x, lam, sigma), ref.hessian(x, lam, sigma), "hess")
        store.allgather(b"")
        if dist is not None:
            dist.destroy_process_group()
        else:
            store.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world,target,kind", [(2, "_worker", "gloo"), (3, "_worker", "gloo"), (2, "_worker_c5", "gloo"),
                                               (2, "_worker", "socket"), (3, "_worker_c5", "socket")])
def test_row_sharded_matches_global_oracle(world, target, kind):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    fn = globals()[target]
    procs = [ctx.Process(target=fn, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def _store_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        from dnlp_b200 import comm
        st = comm.SocketStore(rank, world, "127.0.0.1", port)
        parts = st.allgather(b"r%d" % rank * (rank + 1))
        assert parts == [b"r%d" % r * (r + 1) for r in range(world)]
        root = min(1, world - 1)
        assert comm.bcast(st, b"hello" if rank == root else b"x", root=root) == b"hello"
        big = np.arange(300000, dtype=np.float64) * (rank + 1)            # > one TCP segment
        total = comm.allreduce_sum(st, big)
        np.testing.assert_array_equal(total, np.arange(300000, dtype=np.float64) * sum(range(1, world + 1)))
        np.testing.assert_array_equal(comm.allreduce_max(st, np.array([rank, -rank], float)), [world - 1, 0])
        ragged = comm.allgather_array(st, np.arange(rank, dtype=np.int32))
        assert [a.size for a in ragged] == list(range(world))
        comm.barrier(st)
        st.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world", [1, 2, 4])
def test_socket_store_collectives(world):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_store_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_pair_runs_follow_both_index_maps():
    from dnlp_b200.sharded import _pair_runs
    assert _pair_runs([], []) == []
    assert _pair_runs([0, 1, 2, 3], [10, 11, 12, 13]) == [(0, 10, 4)]
    # a break in either map starts a new run
    assert _pair_runs([0, 1, 2, 5, 6], [10, 11, 12, 13, 14]) == [(0, 10, 3), (5, 13, 2)]
    assert _pair_runs([0, 1, 2, 3], [10, 11, 20, 21]) == [(0, 10, 2), (2, 20, 2)]
    assert _pair_runs(np.arange(0, 200, 2), np.arange(100)) is None           # too fragmented: NVLink route
    assert len(_pair_runs(np.arange(0, 20, 2), np.arange(10))) == 10


# ------------------------------------------------------------------------------------------------
# worker loop, host logic only: a stand-in for the device shard records what each rank is asked to
# evaluate (the real one is csrc/dnlp_shard.cu, exercised in tests/test_sharded_nccl.py on a GPU)
# ------------------------------------------------------------------------------------------------
class _Bus:
    def __init__(self, n, m):
        import queue
        import threading
        self.x, self.lam, self.q = np.full(n, np.nan), np.full(max(m, 1), np.nan), queue.Queue()
        self.both = threading.Barrier(2, timeout=20)     # every evaluation is collective: nobody runs ahead


class _FakeDeviceShard:
    """The attributes and methods RowShardedOracles uses on its device shard.  ``eval`` keeps the vectors it was
    handed (None = "what you have") and logs (callback, x was resent, lambda was resent)."""
    global_inputs, shared = True, {}

    def __init__(self, bus, is_root, m):
        self.bus, self.is_root, self.m = bus, is_root, m
        self.serving, self._force_post = False, 3
        self.x_in, self.lam_in = bus.x, bus.lam
        self.staged_x, self.staged_lam, self.log = None, None, []

    def post(self, name, x, lam=None, sigma=1.0):           # what dnlp_shard_post_command does
        if name is None:
            self.bus.q.put((None, 0.0, 0))
            return None, None
        flags = 0
        if (self._force_post & 1) or not np.array_equal(self.bus.x, x):
            self.bus.x[:] = x
            flags |= 1
        if lam is not None and ((self._force_post & 2) or not np.array_equal(self.bus.lam[:self.m], lam[:self.m])):
            self.bus.lam[:self.m] = lam[:self.m]
            flags |= 2
        self._force_post &= ~(3 if (name == "hess" and lam is not None) else 1)
        self.bus.q.put((name, float(sigma), flags))
        return (self.x_in if flags & 1 else None), (self.lam_in if (lam is not None and flags & 2) else None)

    def wait(self, timeout_s=1.0):
        import queue
        try:
            return self.bus.q.get(timeout=timeout_s)
        except queue.Empty:
            return False, 0.0, 0

    def eval(self, name, x, lam=None, sigma=1.0):
        if x is not None:
            self.staged_x = np.array(x, copy=True)
        if lam is not None:
            self.staged_lam = np.array(lam[:self.m], copy=True)
        assert self.staged_x is not None, "asked to keep a point that was never staged"
        if name == "hess":
            assert self.staged_lam is not None, "asked to keep multipliers that were never staged"
        self.log.append((name, x is not None, lam is not None, float(sigma), self.staged_x.copy(),
                         None if self.staged_lam is None else self.staged_lam.copy()))
        if self.serving:
            self.bus.both.wait()
        return np.float64(0.0) if name == "f" else np.zeros(1)


def test_worker_loop_host_logic_with_a_recording_device():
    """Root posts, worker follows: same callbacks in the same order on the same (x, lambda, sigma); a vector the
    root found unchanged is resent to nobody; switching to SPMD calls and back re-stages both vectors."""
    import threading

    from dnlp_b200 import workloads as W
    from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_logistic_regression
    from oracle.dnlp_oracle import RefOracles
    At, x_init = W.logistic_data(60, 6, 3, seed=5)
    glob = W.logistic_regression(At, x_init)
    gs = GlobalStructure.from_problem(glob)
    bus = _Bus(glob.n, glob.m)
    ranks = []
    for r in range(2):
        local, layout = shard_logistic_regression(At, x_init, r, 2)
        o = RowShardedOracles(local, layout, gs, oracle_factory=RefOracles)
        o._dev = _FakeDeviceShard(bus, r == 0, glob.m)
        assert o.has_worker_loop
        ranks.append(o)
    root, worker = ranks
    with pytest.raises(RuntimeError):
        root.set_worker_loop(True) or root.serve()           # the root runs the solver
    for o in ranks:
        o.set_worker_loop(True)
    served = []
    t = threading.Thread(target=lambda: served.append(worker.serve(poll_s=0.05)))
    t.start()
    rng = np.random.default_rng(0)
    x1, x2 = glob.x0 * 1.01, glob.x0 * 0.97
    lam1, lam2 = rng.standard_normal(glob.m), rng.standard_normal(glob.m + glob.n)    # Knitro hands over m + n duals
    root.objective(x1), root.gradient(x1.copy()), root.constraints(list(x1)), root.jacobian(x1)
    root.hessian(x1, lam1, 1.0), root.hessian(x1, lam1, 0.5), root.hessian(x1, lam2, 0.5)
    root.objective(x2), root.hessian(x2, lam2, 0.0)
    root.release_workers()
    t.join(timeout=20)
    assert served == [9]
    want = [("f", True, False), ("grad", False, False), ("g", False, False), ("jac", False, False),
            ("hess", False, True), ("hess", False, False), ("hess", False, True), ("f", True, False),
            ("hess", False, False)]
    for dev in (root._dev, worker._dev):
        assert [(e[0], e[1], e[2]) for e in dev.log] == want
    for a, b in zip(root._dev.log, worker._dev.log):          # both ranks evaluated the same point every time
        assert a[3] == b[3]
        np.testing.assert_array_equal(a[4], b[4])
        if a[0] == "hess":
            np.testing.assert_array_equal(a[5], b[5])
    np.testing.assert_array_equal(root._dev.log[-1][4], x2)
    np.testing.assert_array_equal(root._dev.log[-1][5], lam2[:glob.m])
    assert [e[3] for e in root._dev.log if e[0] == "hess"] == [1.0, 0.5, 0.5, 0.0]
    # SPMD calls in between move the devices' point; back in the loop the first posts must re-stage x AND lambda
    for o in ranks:
        o.set_worker_loop(False)
        o.objective(glob.x0), o.hessian(glob.x0, lam1, 2.0)
        assert o._dev.log[-1][:3] == ("hess", True, True)
        o.set_worker_loop(True)
    t = threading.Thread(target=lambda: served.append(worker.serve(poll_s=0.05)))
    t.start()
    root.objective(x2), root.hessian(x2, lam2, 0.0)           # the bus still holds x2 / lam2: the compare says "same"
    root.release_workers()
    t.join(timeout=20)
    assert served == [9, 2]
    for dev in (root._dev, worker._dev):
        assert [(e[0], e[1], e[2]) for e in dev.log[-2:]] == [("f", True, False), ("hess", False, True)]
        np.testing.assert_array_equal(dev.log[-1][4], x2)
        np.testing.assert_array_equal(dev.log[-1][5], lam2[:glob.m])


# ------------------------------------------------------------------------------------------------
# host assembly over many layouts at once: ranks as threads of this process (no rendezvous cost)
# ------------------------------------------------------------------------------------------------
class _ThreadStore:
    """``allgather`` among the threads of one process: the store interface without sockets."""

    def __init__(self, world):
        import threading
        self.world, self.bar, self.slots = world, threading.Barrier(world, timeout=60), [None] * world

    def view(self, rank):
        import types
        st = self

        def allgather(payload):
            st.slots[rank] = bytes(payload)
            st.bar.wait()
            out = list(st.slots)
            st.bar.wait()
            return out
        return types.SimpleNamespace(rank=rank, world=st.world, allgather=allgather)


@pytest.mark.parametrize("seed", range(6))
def test_row_sharded_assembly_over_random_layouts_including_empty_shards(seed):
    """World sizes 1..6, both shard builders, row counts down to ONE (so some ranks own no rows at all): every rank's
    five outputs equal the global oracle's (400 such cases were run offline; six seeds x four cases stay here)."""
    import threading
    import traceback

    from dnlp_b200 import workloads as W
    from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_logistic_regression, shard_microbench
    from golden_util import assert_close
    from oracle.dnlp_oracle import RefOracles
    rng = np.random.default_rng(100 + seed)
    for case in range(4):
        kind = "c3" if (seed + case) % 2 == 0 else "c5"
        world = int(rng.integers(1, 7))
        m = 1 if case == 0 else int(rng.integers(2, 30))          # case 0: fewer rows than ranks
        if kind == "c3":
            n = int(rng.integers(2, 10))
            At, x_init = W.logistic_data(m, n, int(rng.integers(1, min(n, 4) + 1)), seed=int(rng.integers(0, 10 ** 6)))
            glob = W.logistic_regression(At, x_init)
            make = lambda r: shard_logistic_regression(At, x_init, r, world)          # noqa: E731
        else:
            A, x0 = W.microbench_data(8 * int(rng.integers(1, 4)), m, int(rng.integers(1, 4)), seed=int(rng.integers(0, 10 ** 6)))
            glob = W.microbench(A, x0)
            make = lambda r: shard_microbench(A, x0, r, world)                       # noqa: E731
        gs = GlobalStructure.from_problem(glob)
        ref = RefOracles(glob)
        ref.jacobianstructure(), ref.hessianstructure()
        x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
        lam = rng.standard_normal(glob.m)
        want = {"f": ref.objective(x), "grad": ref.gradient(x).copy(), "g": np.array(ref.constraints(x)),
                "jac": np.array(ref.jacobian(x)), "hess": np.array(ref.hessian(x, lam, 0.7))}
        st, errs = _ThreadStore(world), []

        def worker(r):
            try:
                local, layout = make(r)
                o = RowShardedOracles(local, layout, gs, store=st.view(r), oracle_factory=RefOracles)
                got = {"f": o.objective(x), "grad": o.gradient(x), "g": o.constraints(x), "jac": o.jacobian(x),
                       "hess": o.hessian(x, lam, 0.7)}
                for key in got:
                    assert_close(got[key], want[key], "%s on rank %d of %d (%s, m=%d)" % (key, r, world, kind, m), atol=1e-11)
            except Exception:           # noqa: BLE001
                errs.append(traceback.format_exc())
                st.bar.abort()
        threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        [t.start() for t in threads]
        [t.join(timeout=120) for t in threads]
        assert not errs, errs[0]
