"""Scheduling metadata of the tape, checked on the CPU for every golden problem.

The CUDA engine trusts three things the compiler writes into each instruction:
  * ``deps``      - which instructions produce the V ranges it reads.  Independent instructions are
                    replayed as parallel CUDA-graph branches, so a missing entry would be a data race;
  * ``dep_mask``  - whether the result depends (transitively) on x, sigma, lambda.  Results stay valid
                    until one of those inputs changes, so a missing bit would serve stale numbers;
  * ``dynamic_sigma`` - Hessian entries that depend on sigma only (kept on the host between calls).
They are re-derived here independently from the instruction operands, and the engine's caching policy
(dnlp_cabi.cu: cacheable / invalidate / overwrite-only output layers) is replayed with the NumPy tape
interpreter over random interleavings of new x, new lambda and new sigma.
"""
import numpy as np
import pytest

from dnlp_b200 import tape as T
from dnlp_b200.compiler import compile_problem
from golden_util import Golden, golden_names
from tape_interp import TapeInterp


def _read_slots(ins):
    if ins.kind == T.K_ELEM:
        r = [ins.a_off + np.arange(ins.count) * ins.a_stride]
        if ins.fcode in T.BINARY_CODES:
            r.append(ins.b_off + np.arange(ins.count) * ins.b_stride)
        return np.unique(np.concatenate(r))
    if ins.kind == T.K_POLY:
        s = np.concatenate([ins.f1, ins.f2])
        return np.unique(s[s >= 0])
    if ins.kind == T.K_SPMVJ:
        s = np.concatenate([ins.f1, ins.f1[ins.qpos >= 0] + 1])
        return np.unique(s[s >= 0])
    if ins.kind == T.K_GEMV:
        return np.arange(ins.x_off, ins.x_off + ins.ncols)
    if ins.kind == T.K_SCALE:
        return np.array([ins.s_slot])
    raise AssertionError(ins.kind)


def _derive(tape):
    """(deps, mask) per instruction, from the operands alone."""
    n, m = tape.n, tape.m
    producer = np.full(tape.nslots, -1, dtype=np.int64)
    deps, masks = [], []
    for ins in tape.instrs:
        slots = _read_slots(ins)
        d = set(int(p) for p in np.unique(producer[slots[slots >= n + 1 + m]]) if p >= 0)
        mk = ((T.DEP_X if np.any(slots < n) else 0) | (T.DEP_SIGMA if np.any(slots == n) else 0)
              | (T.DEP_LAMBDA if np.any((slots > n) & (slots < n + 1 + m)) else 0))
        for p in d:
            mk |= masks[p]
        deps.append(d)
        masks.append(mk)
        if ins.dst_space == T.DST_V:
            st = getattr(ins, "dst_stride", 1) if ins.kind == T.K_ELEM else 1
            w = ins.dst_off + st * np.arange(ins.count)
            assert np.all(producer[w] == -1), "a V slot is written twice"
            producer[w] = ins.id
    # nothing reads a temporary before it is produced
    for ins in tape.instrs:
        slots = _read_slots(ins)
        tmp = slots[slots >= n + 1 + m]
        assert np.all(producer[tmp] >= 0) and np.all(producer[tmp] < ins.id)
    return deps, masks


def _knobs(monkeypatch, layered):
    if layered:
        from dnlp_b200.rules import Builder
        monkeypatch.setattr(Builder, "LAYER_MIN", 2)
        monkeypatch.setattr(Builder, "LONG_ROW", 3)
        monkeypatch.setattr(Builder, "CHUNK", 2)


@pytest.mark.parametrize("layered", [False, True])
@pytest.mark.parametrize("name", golden_names())
def test_deps_and_dep_mask_match_the_operands(name, layered, monkeypatch):
    _knobs(monkeypatch, layered)
    tape = compile_problem(Golden(name).problem)
    deps, masks = _derive(tape)
    for ins in tape.instrs:
        assert set(ins.deps) == deps[ins.id], "instr %d: deps %s, operands say %s" % (ins.id, ins.deps, deps[ins.id])
        assert all(d < ins.id for d in ins.deps)
        assert ins.dep_mask == masks[ins.id], "instr %d: dep_mask %d, operands say %d" % (ins.id, ins.dep_mask, masks[ins.id])
        assert bool(ins.uses_lam) == bool(ins.dep_mask & (T.DEP_SIGMA | T.DEP_LAMBDA))
    # every program is closed under deps and topologically ordered
    for pname, prog in tape.programs.items():
        seen = set()
        for i in prog:
            assert set(tape.instrs[i].deps) <= seen, "program %s runs %d before its inputs" % (pname, i)
            seen.add(i)


class EngineModel(TapeInterp):
    """dnlp_cabi.cu's validity policy on top of the NumPy interpreter: persistent output arrays,
    instructions skipped while valid, flags cleared by what changed."""

    def __init__(self, tape):
        super().__init__(tape)
        t = tape
        self.outs = {T.DST_F: np.array([t.f_const]), T.DST_GRAD: t.grad_const.copy(), T.DST_G: t.g_const.copy(),
                     T.DST_JAC: t.jac_const.copy(), T.DST_HESS: t.hess_const.copy()}
        self.valid = np.zeros(len(t.instrs), dtype=bool)
        self.space_has_acc = {s: any(i.accumulate and i.dst_space == s for i in t.instrs) for s in range(1, 6)}
        self.x = self.lam = self.sigma = None
        self.skipped = 0

    def cacheable(self, ins):
        if ins.dst_space == T.DST_V:
            return not ins.uses_lam
        return ins.dep_mask == T.DEP_SIGMA and not ins.accumulate and not self.space_has_acc[ins.dst_space]

    def invalidate(self, bits):
        for ins in self.t.instrs:
            if ins.dep_mask == 0 or (ins.dep_mask & bits):
                self.valid[ins.id] = False

    def call(self, name, x, lam=None, sigma=None):
        t = self.t
        if self.x is None or not np.array_equal(x, self.x):
            self.V[:t.n] = x
            self.x = x.copy()
            self.invalidate(T.DEP_X)
        if lam is not None:
            if self.sigma is None or sigma != self.sigma:
                self.invalidate(T.DEP_SIGMA)
                self.sigma = sigma
                self.V[t.n] = sigma
            if self.lam is None or not np.array_equal(lam, self.lam):
                self.invalidate(T.DEP_LAMBDA)
                self.lam = lam.copy()
                self.V[t.n + 1:t.n + 1 + t.m] = lam
        plan = []
        for i in t.programs[name]:
            ins = t.instrs[i]
            if self.cacheable(ins) and self.valid[i]:
                self.skipped += 1
                continue
            plan.append(i)
        self._run(plan, self.outs)
        for i in plan:
            if self.cacheable(t.instrs[i]):
                self.valid[i] = True
        return self.outs[{"f": T.DST_F, "grad": T.DST_GRAD, "g": T.DST_G, "jac": T.DST_JAC, "hess": T.DST_HESS}[name]]


@pytest.mark.parametrize("layered", [False, True])
@pytest.mark.parametrize("name", golden_names())
def test_validity_policy_never_serves_stale_values(name, layered, monkeypatch):
    _knobs(monkeypatch, layered)
    g = Golden(name)
    tape = compile_problem(g.problem)
    eng, fresh = EngineModel(tape), TapeInterp(tape)
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 32)
    x0 = g.points[0]["x"]
    xs = [x0 * (1 + 0.02 * rng.standard_normal(x0.size)) + 0.01 * rng.standard_normal(x0.size) for _ in range(3)]
    lams = [rng.standard_normal(tape.m) for _ in range(3)]
    sigmas = [1.0, 1.0, 0.5, 0.0]
    with np.errstate(all="ignore"):
        for step in range(40):
            x, lam, sg = xs[rng.integers(3)], lams[rng.integers(3)], sigmas[rng.integers(4)]
            which = ("f", "grad", "g", "jac", "hess")[rng.integers(5)]
            if which == "hess":
                got, want = eng.call("hess", x, lam, sg), fresh.eval("hess", x, lam, sg)
            else:
                got, want = eng.call(which, x), fresh.eval(which, x)
            np.testing.assert_array_equal(got, want, err_msg="%s at step %d" % (which, step))
    if name in ("c2_eigen_qcqp_small", "c3_logistic_small", "c5_microbench_small"):
        assert eng.skipped > 0                     # the policy does skip work on the BASELINE shapes


@pytest.mark.parametrize("name", golden_names())
def test_sigma_only_hessian_entries(name):
    """``dynamic_sigma``: unchanged under new x and lambda, proportional to sigma, disjoint from the
    compile-time constants; everything outside ``dynamic`` never changes."""
    g = Golden(name)
    tape = compile_problem(g.problem)
    dyn, sig = tape.dynamic[T.DST_HESS], tape.dynamic_sigma[T.DST_HESS]
    assert np.all(np.isin(sig, dyn))
    it = TapeInterp(tape)
    rng = np.random.default_rng(5)
    x0 = g.points[0]["x"]
    x1 = x0 * (1 + 0.05 * rng.standard_normal(x0.size))
    l0, l1 = rng.standard_normal(tape.m), rng.standard_normal(tape.m)
    with np.errstate(all="ignore"):
        a = it.eval("hess", x0, l0, 0.7).copy()
        b = it.eval("hess", x1, l1, 0.7).copy()
        c = it.eval("hess", x1, l1, 1.4).copy()
    const = np.setdiff1d(np.arange(a.size), dyn)
    np.testing.assert_array_equal(a[const], tape.hess_const[const])
    np.testing.assert_array_equal(b[const], a[const])
    np.testing.assert_array_equal(a[sig], b[sig])
    np.testing.assert_allclose(c[sig], 2.0 * b[sig], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", ["c5_microbench_small", "matmul_with_atom_operand", "matmul_const_sides", "c5_lifted_small"])
def test_metadata_with_interleaved_pairs_and_fused_instruction(name, monkeypatch):
    """Pair regions have two writers (even slots: value, odd slots: derivative) and the fused SPMVJ
    instruction reads both: deps / dep_mask must still be exactly what the operands say."""
    from dnlp_b200.rules import Builder
    monkeypatch.setattr(Builder, "PAIRING", True)        # off by default (measured slower at the C5 size)
    monkeypatch.setattr(Builder, "PAIR_MIN_NNZ", 1)
    tape = compile_problem(Golden(name).problem)
    deps, masks = _derive(tape)
    for ins in tape.instrs:
        assert set(ins.deps) == deps[ins.id], "instr %d: deps %s, operands say %s" % (ins.id, ins.deps, deps[ins.id])
        assert ins.dep_mask == masks[ins.id]
