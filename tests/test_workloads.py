"""The native IR builders for the BASELINE configs reproduce what the reference's own reduction
chain emits (checked against the golden fixtures, which were produced by the live reference)."""
import numpy as np
import scipy.sparse as sp

from dnlp_b200 import workloads as W
from dnlp_b200.compiler import compile_problem
from golden_util import Golden, assert_close
from tape_interp import TapeInterp


def _check(prob, g):
    tape = compile_problem(prob)
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.jac_cols, g.jac_cols)
    np.testing.assert_array_equal(tape.hess_rows, g.hess_rows)
    np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    it = TapeInterp(tape)
    for p in g.points:
        assert_close(it.eval("f", p["x"]), p["f"], "f")
        assert_close(it.eval("grad", p["x"]), p["grad"], "grad")
        assert_close(it.eval("g", p["x"]), p["g"], "g")
        assert_close(it.eval("jac", p["x"]), p["jac"], "jac")
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")


def test_c1_toy_matches_reference_chain():
    g = Golden("c1_readme_toy")
    prob = W.eigen_qcqp(3)
    np.testing.assert_array_equal(prob.x0, g.problem.x0)
    _check(prob, g)


def test_c2_matches_reference_chain():
    _check(W.eigen_qcqp(24), Golden("c2_eigen_qcqp_small"))


def test_c3_matches_reference_chain():
    rng = np.random.default_rng(0)                      # same draws as tests/golden/make_golden.py
    m, n, k = 200, 16, 4
    cols = np.concatenate([rng.choice(n, k, replace=False) for _ in range(m)])
    A = sp.csr_matrix((rng.standard_normal(m * k), (np.repeat(np.arange(m), k), cols)), shape=(m, n))
    y = rng.choice([-1.0, 1.0], m)
    At = sp.diags(-y) @ A
    x_init = 0.1 * rng.standard_normal(n)
    g = Golden("c3_logistic_small")
    prob = W.logistic_regression(sp.csr_array(At), x_init)
    np.testing.assert_array_equal(prob.lb, g.problem.lb)
    assert_close(prob.x0, g.problem.x0, "x0")
    _check(prob, g)


def test_c4_matches_reference_chain():
    P, q, rng = W.qcqp_data(12, 3)
    _check(W.qcqp(P, q), Golden("c4_qcqp_small"))


def test_c5_matches_reference_chain():
    rng = np.random.default_rng(0)
    seg, m, k = 5, 20, 10
    N = seg * 8
    x0 = np.concatenate([rng.uniform(0.5, 1.5, seg) for _ in range(8)])
    cols = np.concatenate([rng.choice(N, k, replace=False) for _ in range(m)])
    A = sp.csr_matrix((rng.standard_normal(m * k), (np.repeat(np.arange(m), k), cols)), shape=(m, N))
    g = Golden("c5_microbench_small")
    prob = W.microbench(sp.csr_array(A), x0)
    np.testing.assert_array_equal(prob.x0, g.problem.x0)
    _check(prob, g)


def test_distinct_columns_are_distinct():
    cols = W.distinct_columns(np.random.default_rng(3), 500, 40, 16)
    assert cols.shape == (500, 16)
    assert all(len(set(r)) == 16 for r in cols)


# ---- fused constraint value + Jacobian fill (SPMVJ) and the interleaved pair layout ---------------------
def _union(tape, x, lam, sigma):
    """Run the union program with the NumPy interpreter; returns the five outputs."""
    from dnlp_b200 import tape as T
    it = TapeInterp(tape)
    it.V[:tape.n] = x
    it.V[tape.n] = sigma
    it.V[tape.n + 1:tape.n + 1 + tape.m] = lam
    outs = {T.DST_F: np.array([tape.f_const]), T.DST_GRAD: tape.grad_const.copy(), T.DST_G: tape.g_const.copy(),
            T.DST_JAC: tape.jac_const.copy(), T.DST_HESS: tape.hess_const.copy()}
    it._run(tape.programs["all"], outs)
    return outs


def test_spmv_jacobian_fusion_on_c5(monkeypatch):
    """With the pairing threshold lowered the C5 fixture compiles to ONE fused instruction in the union
    program (value slots even, derivative right after), the per-callback programs keep their own kernels,
    and every program still reproduces the live reference."""
    from dnlp_b200 import tape as T
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "PAIRING", True)        # off by default (measured slower at the C5 size)
    monkeypatch.setattr(Builder, "PAIR_MIN_NNZ", 1)
    g = Golden("c5_microbench_small")
    tape = compile_problem(g.problem)
    fused = [i for i in tape.instrs if i.kind == T.K_SPMVJ]
    assert len(fused) == 1
    F = fused[0]
    assert F.id in tape.programs["all"] and F.id not in tape.programs["g"] and F.id not in tape.programs["jac"]
    assert F.fused_jac[0] in tape.programs["g"] and F.fused_jac[1] in tape.programs["jac"]
    assert F.fused_jac[0] not in tape.programs["all"] and F.fused_jac[1] not in tape.programs["all"]
    assert np.all(F.f1[F.f1 >= 0] % 2 == 0)
    assert np.array_equal(np.sort(F.qpos[F.qpos >= 0]), np.arange(tape.jac_rows.size))     # every entry exactly once
    strided = [i for i in tape.instrs if i.kind == T.K_ELEM and i.dst_stride == 2]
    assert len(strided) == 16 and any(i.post_scale != 1.0 for i in strided)                  # 8 atoms x (value, derivative)
    np.testing.assert_array_equal(tape.jac_rows, g.jac_rows)
    np.testing.assert_array_equal(tape.hess_cols, g.hess_cols)
    it = TapeInterp(tape)
    for p in g.points:
        for name in ("f", "grad", "g", "jac"):
            assert_close(it.eval(name, p["x"]), p[name], name)
        assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], "hess")
        outs = _union(tape, p["x"], p["lam"], float(p["sigma"]))
        for space, name in ((T.DST_F, "f"), (T.DST_GRAD, "grad"), (T.DST_G, "g"), (T.DST_JAC, "jac"), (T.DST_HESS, "hess")):
            assert_close(outs[space], p[name], "union/" + name)


def test_fusion_is_refused_when_it_would_be_wrong(monkeypatch):
    """No pair candidates (atoms under a matmul with a NON-constant left operand, lifted forms, small
    matrices) -> no fused instruction and no strided ELEM; problems with other constraint shapes keep
    compiling exactly as before."""
    from dnlp_b200 import tape as T
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "PAIRING", True)        # off by default (measured slower at the C5 size)
    monkeypatch.setattr(Builder, "PAIR_MIN_NNZ", 1)
    for name in ("c3_logistic_small", "c5_lifted_small", "matmul_with_atom_operand", "hs071", "clnlbeam"):
        g = Golden(name)
        tape = compile_problem(g.problem)
        assert not any(i.kind == T.K_SPMVJ for i in tape.instrs) or name == "matmul_with_atom_operand"
        it = TapeInterp(tape)
        p = g.points[0]
        outs = _union(tape, p["x"], p["lam"], float(p["sigma"]))
        assert_close(outs[T.DST_JAC], p["jac"], name + " union/jac")
        assert_close(outs[T.DST_G], p["g"], name + " union/g")
        assert_close(it.eval("jac", p["x"]), p["jac"], name + " jac")


def test_column_panels_split_and_accumulate(monkeypatch):
    """SpMV-shaped outputs whose gathered slots span more than the L2 keeps are cut into column panels:
    pass 1 writes the row sums, pass 2 adds its own (same rows, disjoint terms, every term exactly once)."""
    from dnlp_b200 import tape as T
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "PANEL_MIN_TERMS", 8)
    monkeypatch.setattr(Builder, "PANEL_SPAN_BYTES", 64)
    for name in ("c5_microbench_small", "c3_logistic_small", "matmul_const_sides"):
        g = Golden(name)
        tape = compile_problem(g.problem)
        acc = [i for i in tape.instrs if i.kind == T.K_POLY and i.accumulate]
        if name == "c5_microbench_small":
            assert acc, "no panel pass emitted"
        for i in acc:
            first = tape.instrs[i.panel_prev]
            assert first.dst_space == i.dst_space != T.DST_V and first.count == i.count and not first.accumulate
            assert first.id < i.id
            for prog in tape.programs.values():
                assert (i.id in prog) == (first.id in prog)
            lo = i.f1[i.f1 >= 0].min()
            assert first.f1.max() < lo                       # disjoint slot ranges
        it = TapeInterp(tape)
        for p in g.points:
            assert_close(it.eval("g", p["x"]), p["g"], name + " g")
            assert_close(it.eval("jac", p["x"]), p["jac"], name + " jac")
            assert_close(it.eval("hess", p["x"], p["lam"], float(p["sigma"])), p["hess"], name + " hess")
            outs = _union(tape, p["x"], p["lam"], float(p["sigma"]))
            assert_close(outs[T.DST_G], p["g"], name + " union/g")


def test_kron_fast_path_is_scipys_kron_entry_for_entry():
    """rules.Builder._kron skips SciPy's repeat / tile passes when one side is a 1 x 1 identity (matrix @ vector);
    the entries and their order must be exactly SciPy's, or the triplet order of the reference is lost."""
    import scipy.sparse as sp
    from dnlp_b200.rules import Builder
    rng = np.random.default_rng(0)
    for trial in range(24):
        m, n = rng.integers(1, 9, 2)
        A = sp.random(m, n, density=0.4, random_state=int(rng.integers(1e6)), format=("csr", "coo", "csc")[trial % 3])
        A = sp.coo_array(A) if trial % 2 else A
        for fmt in ("csr", "coo"):
            for left, right in ((sp.eye(1), A), (A.T, sp.eye(1)), (sp.eye(2), A), (A, sp.eye(3))):
                a, b = sp.coo_array(Builder._kron(left, right, fmt)), sp.coo_array(sp.kron(left, right, format=fmt))
                assert a.shape == b.shape
                np.testing.assert_array_equal(a.coords[0], b.coords[0])
                np.testing.assert_array_equal(a.coords[1], b.coords[1])
                np.testing.assert_array_equal(a.data, b.data)


def _same_tapes(a, b, path="tape"):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        assert isinstance(a, np.ndarray) and isinstance(b, np.ndarray), path
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True), path
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same_tapes(x, y, "%s[%d]" % (path, i))
    elif isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same_tapes(a[k], b[k], "%s.%s" % (path, k))
    elif hasattr(a, "__slots__"):
        for k in a.__slots__:
            _same_tapes(getattr(a, k, None), getattr(b, k, None), "%s.%s" % (path, k))
    elif hasattr(a, "__dict__") and not callable(a):
        for k in a.__dict__:
            _same_tapes(a.__dict__[k], b.__dict__.get(k), "%s.%s" % (path, k))
    else:
        assert a == b or (a != a and b != b), (path, a, b)


def test_diagonal_chain_rule_shortcut_emits_the_same_tape(monkeypatch):
    """rules.Builder._chain_through_diagonal writes SciPy's SpGEMM order down directly (reversed first touch per
    row); the generic route runs SciPy on the patterns.  Same tape, entry for entry, incl. rows of A with several
    entries, empty rows and a lifted (log / entr) variant whose inner Jacobians are not all diagonal."""
    from dnlp_b200.compiler import compile_problem
    from dnlp_b200.rules import Builder
    cases = []
    A, x0 = W.microbench_data(4000, 1531, 6, seed=3)
    cases.append(W.microbench(A, x0))
    A2, x2 = W.microbench_data(800, 1200, 3, seed=4)          # more rows than a segment has columns: repeated columns
    cases.append(W.microbench(A2, x2))
    for prob in cases:
        monkeypatch.setattr(Builder, "CHAIN_FASTPATH", True)
        fast = compile_problem(prob)
        monkeypatch.setattr(Builder, "CHAIN_FASTPATH", False)
        generic = compile_problem(prob)
        _same_tapes(fast, generic)
