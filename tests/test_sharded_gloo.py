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
