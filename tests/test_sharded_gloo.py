"""Row-sharded oracle, world_size 2 and 3 over gloo on CPU.

Checks the host logic of dnlp_b200.sharded (index maps into the global reference pattern,
dynamic-entry compaction, packed all-reduce).  The local evaluator is the CPU oracle here (test
infrastructure); on GPUs it is GpuOracles (tests/test_gpu_parity.py covers that evaluator)."""
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


def _worker_c5(rank, world, port, q):
    """C5-type sharding: every variable replicated, Hessian contributions of all ranks summed."""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
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
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), oracle_factory=RefOracles)
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
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
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
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), oracle_factory=RefOracles)
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
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world,target", [(2, "_worker"), (3, "_worker"), (2, "_worker_c5")])
def test_row_sharded_matches_global_oracle(world, target):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    fn = globals()[target]
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
