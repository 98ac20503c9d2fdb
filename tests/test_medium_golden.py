"""Medium-size fixtures from the live reference (tests/golden/make_golden_medium.py): the bench
workloads' own data generators at sizes that reach the production kernels.

CPU tier: the native workload builders reproduce the reference chain's smooth problem (same x0, same
structures bit for bit), and the CPU oracle port reproduces the reference's values (pins the port at
this size).  GPU tier: the CUDA path through the C-ABI against the same vectors."""
import numpy as np
import pytest

from golden_util import assert_close
from workload_cases import MediumGolden, build, c4_starts

CONFIGS = ["c2", "c3", "c4", "c5"]


@pytest.mark.parametrize("config", CONFIGS)
def test_builder_and_oracle_port_match_reference_medium(config):
    from dnlp_b200.compiler import compile_problem
    from oracle.dnlp_oracle import RefOracles
    g = MediumGolden(config)
    prob = build(config)
    assert (prob.n, prob.m) == (g.n, g.m)
    if config != "c4":
        assert_close(prob.x0, g.x0, "x0")            # the chain's initial point (aux variables included)
    tape = compile_problem(prob)
    g.check_structure(tape.jac_rows, tape.jac_cols, tape.hess_rows, tape.hess_cols)
    r = RefOracles(prob)
    jr, jc = r.jacobianstructure()
    hr, hc = r.hessianstructure()
    g.check_structure(np.asarray(jr, np.int32), np.asarray(jc, np.int32), np.asarray(hr, np.int32), np.asarray(hc, np.int32))
    for i, p in enumerate(g.points):
        with np.errstate(all="ignore"):
            assert_close(r.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(r.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            assert_close(r.constraints(p["x"]), p["g"], "g[%d]" % i, atol=1e-9)
            assert_close(r.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(r.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)


@pytest.mark.gpu
@pytest.mark.parametrize("config", CONFIGS)
def test_gpu_matches_reference_medium(config):
    from dnlp_b200.oracles import GpuOracles
    g = MediumGolden(config)
    o = GpuOracles(build(config))
    try:
        g.check_structure(*o.jacobianstructure(), *o.hessianstructure())
        for i, p in enumerate(g.points):
            sg = float(p["sigma"])
            assert_close(o.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i, atol=1e-9)
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], sg), p["hess"], "hess[%d]" % i)
        p = g.points[-1]
        res = o.eval_all(p["x"], p["lam"], float(p["sigma"]))
        for k in ("f", "grad", "g", "jac", "hess"):
            assert_close(res[k], p[k], "eval_all/" + k, atol=1e-9 if k == "g" else 1e-12)
        kernels = {o.instr_kernel(i) for i in range(len(o.tape.instrs))} - {""}
        assert kernels, "no kernel ran"
    finally:
        o.close()


@pytest.mark.gpu
def test_gpu_batched_c4_matches_reference_medium():
    """The batched engine at the benchmark's n = 512, k = 8: the reference's values at 4 of the 4096
    start points, evaluated inside a batch of 64 (the picked starts first)."""
    from dnlp_b200.multistart import BatchedOracles
    g = MediumGolden("c4")
    X = c4_starts()
    B = 64
    picks = [int(b) for b in g.starts]
    rows = picks + [b for b in range(B) if b not in picks][:B - len(picks)]
    Xb = X[rows]
    rng = np.random.default_rng(5)
    LAM, SIG = rng.standard_normal((B, g.m)), np.ones(B)
    for j, p in enumerate(g.points):
        assert np.array_equal(p["x"], Xb[j])
        LAM[j], SIG[j] = p["lam"], float(p["sigma"])
    o = BatchedOracles(build("c4"), B)
    try:
        g.check_structure(*o.jacobianstructure(), *o.hessianstructure())
        res = o.eval(Xb, LAM, SIG)
        for j, p in enumerate(g.points):
            for k in ("f", "grad", "g", "jac", "hess"):
                assert_close(res[k][j], p[k], "start %d %s" % (picks[j], k))
    finally:
        o.close()
