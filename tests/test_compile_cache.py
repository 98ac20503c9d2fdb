"""Compile-once reuse across solves of the same smooth problem (the reference's best_of loop
re-applies its reduction chain per start, cvxpy/problems/problem.py:1249-1275)."""
import numpy as np
import scipy.sparse as sp

from dnlp_b200 import ir
from dnlp_b200.compile_cache import OracleCache, fingerprint
from dnlp_b200.compiler import compile_problem


def _problem(seed=0, scale=1.0, n=6, x0=None, extra=False):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    S = sp.random(4, n, density=0.5, random_state=3, format="csr")
    x = ir.Variable(n)
    t = ir.Variable(4)
    obj = ir.quad_form(x, A.T @ A * scale) + ir.sum(ir.exp(t))
    cons = [ir.matmul(ir.Constant(S), x) + (-1.0) * t, ir.sum(ir.power(x, 2)) + (-1.0)]
    if extra:
        cons.append(ir.sum(ir.logistic(x)) + (-2.0))
    return ir.ProblemIR(obj, cons, x0=np.ones(n + 4) if x0 is None else x0)


def test_fingerprint_ignores_ids_names_and_start_point():
    a, b = _problem(), _problem(x0=np.linspace(0, 1, 10))
    assert a.variables[0].attrs["id"] != b.variables[0].attrs["id"]      # fresh ids, as after a re-applied chain
    assert fingerprint(a) == fingerprint(b)
    ta, tb = compile_problem(a), compile_problem(b)
    np.testing.assert_array_equal(ta.hess_rows, tb.hess_rows)
    assert len(ta.instrs) == len(tb.instrs)


def test_fingerprint_sees_constants_structure_and_layout():
    base = fingerprint(_problem())
    assert fingerprint(_problem(scale=1.0000001)) != base                # one constant differs in the last bits
    assert fingerprint(_problem(seed=1)) != base
    assert fingerprint(_problem(extra=True)) != base
    assert fingerprint(_problem(n=7)) != base
    p = _problem()
    q = ir.ProblemIR(p.objective, p.constraints, variables=list(reversed(p.variables)), x0=p.x0)
    assert fingerprint(q) != base                                        # same tree, other variable layout


class _FakeOracle:
    closed = 0

    def __init__(self, prob):
        self.prob = prob

    def close(self):
        _FakeOracle.closed += 1


def test_oracle_cache_lru_and_eviction():
    _FakeOracle.closed = 0
    c = OracleCache(capacity=2)
    o1, hit = c.get(_problem(), _FakeOracle)
    assert not hit
    o1b, hit = c.get(_problem(x0=np.zeros(10)), _FakeOracle)
    assert hit and o1b is o1
    o2, _ = c.get(_problem(seed=1), _FakeOracle)
    o1c, hit = c.get(_problem(), _FakeOracle)                            # refreshes o1
    assert hit and o1c is o1
    c.get(_problem(seed=2), _FakeOracle)                                 # evicts o2 (least recently used)
    assert _FakeOracle.closed == 0        # evicted oracles are dropped, never closed: a caller may still hold them
    _, hit = c.get(_problem(seed=1), _FakeOracle)
    assert not hit
    assert (c.hits, c.misses) == (2, 4)
    c.clear()
    assert _FakeOracle.closed == 0 and len(c._items) == 0
    # the extra key separates devices / build options
    d = OracleCache(capacity=4)
    a0, _ = d.get(_problem(), _FakeOracle, extra_key=(("device", 0),))
    a1, hit = d.get(_problem(), _FakeOracle, extra_key=(("device", 1),))
    assert not hit and a1 is not a0
    a0b, hit = d.get(_problem(), _FakeOracle, extra_key=(("device", 0),))
    assert hit and a0b is a0
    off = OracleCache(capacity=0)
    a, _ = off.get(_problem(), _FakeOracle)
    b, hit = off.get(_problem(), _FakeOracle)
    assert a is not b and not hit


def test_fingerprint_zero_size_and_sparse_order():
    """a constant with an empty dimension must not crash the digest; two sparse constants
    with the same entries in a different storage order are different tapes (the rules emit triplets in
    the constant's own COO order)."""
    import scipy.sparse as sp
    from dnlp_b200.compile_cache import _feed_value
    import hashlib

    def dig(v):
        h = hashlib.blake2b(digest_size=20)
        _feed_value(h, v)
        return h.hexdigest()
    assert dig(np.zeros((0, 3))) != dig(np.zeros((3, 0)))
    r, c, v = np.array([0, 1, 1]), np.array([1, 0, 2]), np.array([1.0, 2.0, 3.0])
    a = sp.coo_array((v, (r, c)), shape=(2, 3))
    b = sp.coo_array((v[::-1], (r[::-1], c[::-1])), shape=(2, 3))
    assert dig(a) != dig(b)
    assert dig(a) == dig(sp.coo_array((v.copy(), (r.copy(), c.copy())), shape=(2, 3)))
    assert dig(a.tocsr()) != dig(a.tocsc())
