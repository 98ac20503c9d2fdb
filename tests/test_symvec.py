"""Unit tests of the symbolic value algebra (host logic of the DAG compiler)."""
import numpy as np

from dnlp_b200.symvec import NONE, SymVec


def ev(sv, V):
    """Reference evaluation of a SymVec against a value buffer."""
    t = sv.coef.copy()
    m1 = sv.f1 != NONE
    t[m1] *= V[sv.f1[m1]]
    m2 = sv.f2 != NONE
    t[m2] *= V[sv.f2[m2]]
    out = np.zeros(sv.K)
    np.add.at(out, sv.row, t)
    return out


def test_constructors_and_predicates():
    V = np.arange(10.0) + 1
    c = SymVec.const([1.0, 0.0, -2.0])
    s = SymVec.slot_range(3, 4)
    assert c.is_const_mask().all() and not s.is_const_mask().any()
    np.testing.assert_array_equal(c.const_values(), [1.0, 0.0, -2.0])
    np.testing.assert_array_equal(ev(s, V), V[3:7])
    assert s.contiguous_start() == 3 and SymVec.slots([5, 2]).contiguous_start() is None
    assert s.is_unit() and c.is_unit()
    assert SymVec.zeros(3).K == 3 and SymVec.zeros(3).nterms == 0


def test_linear_ops_match_numpy():
    rng = np.random.default_rng(0)
    V = rng.standard_normal(12)
    a = SymVec.slots([0, 3, 5, 7]).scale([2.0, -1.0, 0.5, 3.0])
    b = SymVec.const([1.0, 2.0, 3.0, 4.0])
    np.testing.assert_allclose(ev(a.add(b), V), ev(a, V) + ev(b, V))
    np.testing.assert_allclose(ev(a.neg(), V), -ev(a, V))
    idx = np.array([3, 3, 0, 1])
    np.testing.assert_allclose(ev(a.gather(idx), V), ev(a, V)[idx])
    grp = np.array([1, 0, 1, 1])
    np.testing.assert_allclose(ev(a.group_sum(grp, 2), V), np.bincount(grp, weights=ev(a, V), minlength=2))
    np.testing.assert_allclose(ev(a.sum_all(), V), [ev(a, V).sum()])
    sc = a.scatter_into(6, [5, 0, 2, 3])
    want = np.zeros(6)
    want[[5, 0, 2, 3]] = ev(a, V)
    np.testing.assert_allclose(ev(sc, V), want)
    cat = SymVec.concat([a, b])
    np.testing.assert_allclose(ev(cat, V), np.concatenate([ev(a, V), ev(b, V)]))
    M = rng.standard_normal((3, 4))
    r, c = np.nonzero(M)
    np.testing.assert_allclose(ev(a.linear_map(r, c, M[r, c], 3), V), M @ ev(a, V))


def test_gather_of_multi_term_entries():
    V = np.arange(8.0) + 1
    a = SymVec.slots([0, 1, 2]).add(SymVec.slots([3, 4, 5])).add(SymVec.const([1, 1, 1]))   # 3 terms per entry
    assert not a.is_unit() and (a.term_counts() == 3).all()
    np.testing.assert_allclose(ev(a.gather([2, 0, 2]), V), ev(a, V)[[2, 0, 2]])


def test_products():
    V = np.arange(8.0) + 1
    a = SymVec.slots([0, 1, 2]).scale([2, 3, 4])
    b = SymVec.slots([5, 6, 7])
    assert a.can_multiply_directly(b)
    np.testing.assert_allclose(ev(a.mul_simple(b), V), ev(a, V) * ev(b, V))
    c = SymVec.const([2.0, 0.5, -1.0])
    np.testing.assert_allclose(ev(a.mul_simple(b).mul_simple(c), V), ev(a, V) * ev(b, V) * ev(c, V))
    assert not a.mul_simple(b).can_multiply_directly(b)          # would need three factors
    assert not a.add(b).can_multiply_directly(b)                 # two terms per entry


def test_simplify_merges_like_terms_and_drops_zeros():
    V = np.arange(6.0) + 1
    a = SymVec.slots([0, 1]).add(SymVec.slots([0, 1]).scale([2.0, -1.0])).add(SymVec.const([0.0, 5.0]))
    s = a.simplify()
    np.testing.assert_allclose(ev(s, V), ev(a, V))
    assert s.nterms == 2                   # entry 0: 3*V0 ; entry 1: 0*V1 dropped, const 5 kept
    z = SymVec.slots([2, 3]).add(SymVec.const([0.0, 0.0])).drop_zero_constants()
    assert z.nterms == 2 and z.is_unit()


def test_stable_order_counting_sort_equals_stable_argsort():
    """Large inputs take the SciPy counting-sort path; it must be the same permutation."""
    from dnlp_b200.symvec import stable_order
    rng = np.random.default_rng(0)
    for n, K in ((70_000, 10), (200_000, 50_000), (100_000, 399_999), (65_536, 1), (1000, 7)):
        key = rng.integers(0, K, n)
        np.testing.assert_array_equal(stable_order(key, K), np.argsort(key, kind="stable"))
        np.testing.assert_array_equal(stable_order(key.astype(np.int32), K), np.argsort(key, kind="stable"))


def test_coo_sum_duplicates_radix_path_equals_key_sort():
    from dnlp_b200.rules import Builder
    from dnlp_b200.symvec import SymVec
    rng = np.random.default_rng(1)
    n, nr, nc = 150_000, 3000, 2500
    rows, cols = rng.integers(0, nr, n), rng.integers(0, nc, n)
    sv = SymVec(n, np.arange(n), rng.standard_normal(n), rng.integers(0, 50, n), np.full(n, -1))
    r, c, out = Builder._coo_sum_duplicates(rows, cols, sv)
    key = rows * nc + cols
    uniq = np.unique(key)
    np.testing.assert_array_equal(r * nc + c, uniq)
    # same sums, same order of the terms inside every merged entry (stable)
    order = np.argsort(key, kind="stable")
    grp = np.searchsorted(uniq, key[order])
    np.testing.assert_array_equal(out.row, grp)
    np.testing.assert_array_equal(out.coef, sv.coef[order])
    np.testing.assert_array_equal(out.f1, sv.f1[order])


def test_merge_of_sorted_runs_is_the_stable_argsort(monkeypatch):
    """``Builder._merge_sorted_runs``: keys made of a few non-decreasing runs (one block per atom) are merged instead
    of sorted; must be the stable permutation, ties between runs resolved in favour of the earlier run."""
    from dnlp_b200.rules import Builder
    rng = np.random.default_rng(2)
    for trial in range(60):
        nruns = int(rng.integers(1, 8))
        sizes = rng.integers(1, 6000, nruns) if trial % 3 else np.r_[rng.integers(5000, 9000), rng.integers(1, 40, nruns - 1)]
        span = int(rng.choice([5, 300, 10 ** 6]))                      # many ties ... none
        key = np.concatenate([np.sort(rng.integers(0, span, int(s))) for s in sizes]).astype(np.int64)
        got = Builder._merge_sorted_runs(key)
        if got is None:
            assert key.size < (1 << 12) or np.count_nonzero(key[1:] < key[:-1]) >= Builder.MERGE_MAX_RUNS
            continue
        np.testing.assert_array_equal(got, np.argsort(key, kind="stable"))
    many = np.tile(np.arange(10, dtype=np.int64), 1000)                # 1000 runs: not this routine's job
    assert Builder._merge_sorted_runs(many) is None
    np.testing.assert_array_equal(Builder._merge_sorted_runs(np.arange(5000, dtype=np.int64)), np.arange(5000))


def test_coo_sum_duplicates_merge_path_equals_key_sort():
    """A long sorted block followed by short ones (dense quad_form Hessian + diagonal blocks) through the merge path:
    same entries, same order of the terms inside every merged entry as the key sort."""
    from dnlp_b200.rules import Builder
    rng = np.random.default_rng(3)
    n = 90
    R, C = np.divmod(np.arange(n * n), n)                              # dense block, row-major
    d = np.arange(n)
    rows, cols = np.concatenate([R, d, d[::2]]), np.concatenate([C, d, d[::2]])
    N = rows.size
    sv = SymVec(N, np.arange(N), rng.standard_normal(N), rng.integers(0, 50, N), np.full(N, -1))
    r, c, out = Builder._coo_sum_duplicates(rows, cols, sv)
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    uniq = np.unique(key)
    np.testing.assert_array_equal(r * n + c, uniq)
    np.testing.assert_array_equal(out.row, np.searchsorted(uniq, key[order]))
    np.testing.assert_array_equal(out.coef, sv.coef[order])
    np.testing.assert_array_equal(out.f1, sv.f1[order])


def test_concat_and_add_many_equal_their_pairwise_forms():
    rng = np.random.default_rng(4)

    def rand(K, nt):
        return SymVec(K, np.sort(rng.integers(0, K, nt)), rng.standard_normal(nt), rng.integers(-1, 9, nt), rng.integers(-1, 9, nt))
    parts = [rand(7, 12), rand(3, 0), rand(5, 9), rand(1, 1)]
    cat = SymVec.concat(parts)
    assert cat.K == 16 and cat.row.dtype == np.int32
    np.testing.assert_array_equal(cat.row, np.concatenate([p.row.astype(np.int64) + o for p, o in zip(parts, (0, 7, 10, 15))]))
    np.testing.assert_array_equal(cat.coef, np.concatenate([p.coef for p in parts]))
    same = [rand(6, 10), rand(6, 4), rand(6, 0), rand(6, 13)]
    chain = same[0].add(same[1]).add(same[2]).add(same[3])
    many = SymVec.add_many(same)
    for name in ("row", "coef", "f1", "f2"):
        np.testing.assert_array_equal(getattr(many, name), getattr(chain, name))
