"""Pin the CPU oracle (oracle/dnlp_oracle.py) against the live-reference golden vectors."""
import numpy as np
import pytest

from golden_util import Golden, assert_close, golden_names
from oracle.dnlp_oracle import RefOracles


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g = Golden(name)
    o = RefOracles(g.problem)
    jr, jc = o.jacobianstructure()
    hr, hc = o.hessianstructure()
    # structures: bit-exact, including order
    assert jr.dtype == np.int32 and hr.dtype == np.int32
    np.testing.assert_array_equal(jr, g.jac_rows)
    np.testing.assert_array_equal(jc, g.jac_cols)
    np.testing.assert_array_equal(hr, g.hess_rows)
    np.testing.assert_array_equal(hc, g.hess_cols)
    with np.errstate(all="ignore"):
        for i, p in enumerate(g.points):
            assert_close(o.objective(p["x"]), p["f"], "f[%d]" % i)
            assert_close(o.gradient(p["x"]), p["grad"], "grad[%d]" % i)
            if g.problem.m:
                assert_close(o.constraints(p["x"]), p["g"], "g[%d]" % i)
            assert_close(o.jacobian(p["x"]), p["jac"], "jac[%d]" % i)
            assert_close(o.hessian(p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess[%d]" % i)
