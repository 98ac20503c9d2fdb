"""Row-sharded oracle over NCCL with device-side assembly (needs >= 2 GPUs; skipped otherwise).
Every rank's assembled global outputs must equal the single-GPU oracle of the global problem."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from dnlp_b200 import workloads as W
        from dnlp_b200.oracles import GpuOracles
        from dnlp_b200.sharded import GlobalStructure, RowShardedOracles, shard_logistic_regression
        from golden_util import assert_close

        At, x_init = W.logistic_data(20011, 48, 8, seed=5)
        glob = W.logistic_regression(At, x_init)
        ref = GpuOracles(glob, device=rank)
        local, layout = shard_logistic_regression(At, x_init, rank, world)
        o = RowShardedOracles(local, layout, GlobalStructure.from_problem(glob), device=rank)
        assert o._devasm is not None, "device-side assembly not active"
        np.testing.assert_array_equal(o.jacobianstructure()[0], ref.jacobianstructure()[0])
        np.testing.assert_array_equal(o.hessianstructure()[1], ref.hessianstructure()[1])
        rng = np.random.default_rng(11)
        for _ in range(3):
            x = glob.x0 * (1 + 0.05 * rng.standard_normal(glob.n))
            lam = rng.standard_normal(glob.m)
            sigma = float(rng.uniform(0.5, 1.5))
            assert_close(o.objective(x), ref.objective(x), "f")
            assert_close(o.gradient(x), ref.gradient(x), "grad")
            assert_close(o.constraints(x), ref.constraints(x), "g", atol=1e-11)
            assert_close(o.jacobian(x), ref.jacobian(x), "jac")
            assert_close(o.hessian(x, lam, sigma), ref.hessian(x, lam, sigma), "hess")
        dist.barrier()
        o.close(), ref.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def test_row_sharded_nccl_device_assembly():
    from dnlp_b200 import _cabi
    if _cabi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
