"""DAG compiler: smooth problem IR -> flat tape + fixed triplet patterns.

Runs once per ``NLPsolver.apply`` (where the reference builds its ``Oracles``
object, reductions/solvers/nlp_solvers/nlp_solver.py:61-79).  It replays, on
symbolic values, exactly what the reference's seven callbacks do numerically on
every call:

  objective      nlp_solver.py:212-216
  gradient       nlp_solver.py:218-235   scatter-ASSIGN of the objective Jacobian
  constraints    nlp_solver.py:237-244
  jacobian       nlp_solver.py:246-307   per constraint, per variable in main_var order;
                                         duplicate-summed CSR sampling when any constraint
                                         is non-affine (insert_missing_zeros_jacobian)
  hessian        nlp_solver.py:337-421   sum_coo (row-major sort + duplicate sum), then the
                                         lower triangle of the structure is sampled

The structures come out bit-identical to the reference's NaN-evaluation passes
(nlp_solver.py:309-335, 374-392) because the index arithmetic is the same and the
SciPy-routed orderings are produced by the same SciPy calls (see rules.py).
"""
import numpy as np

from . import tape as T
from .rules import Builder
from .symvec import NONE, SymVec


def _closure(tape, root_ids):
    need, stack = set(), list(root_ids)
    while stack:
        i = stack.pop()
        if i in need:
            continue
        need.add(i)
        stack.extend(tape.instrs[i].deps)
        if tape.instrs[i].panel_prev is not None:
            stack.append(tape.instrs[i].panel_prev)
    return sorted(need)


SIMPLIFY_LIMIT = 4_000_000     # like-term merging is an optimisation; skipped for huge vectors


def _emit_output(b, sv, space):
    """Split an output vector into its compile-time constant part and the instruction(s) that
    fill the entries depending on x / lambda.  Returns (const_values, [root instructions])."""
    if sv.nterms <= SIMPLIFY_LIMIT:
        sv = sv.simplify()
    else:
        sv = sv.drop_zero_constants()
    cmask = sv.is_const_mask()
    const = np.where(cmask, sv.const_values(), 0.0)
    dyn = np.where(~cmask)[0]
    b.tape.dynamic[space] = dyn.astype(np.int32)
    # entries whose every factor is the objective factor: they change only when sigma does, so a
    # caller that keeps its output array between calls can skip them while sigma is unchanged
    sig = b.tape.sigma_slot
    other = np.zeros(sv.K, dtype=bool)
    if sv.row.size:
        other[sv.row[((sv.f1 != NONE) & (sv.f1 != sig)) | ((sv.f2 != NONE) & (sv.f2 != sig))]] = True
    b.tape.dynamic_sigma[space] = np.where(~cmask & ~other)[0].astype(np.int32)
    if dyn.size == 0:
        return const, []
    if dyn.size == sv.K:
        return const, b.emit_output(sv, space, pos=None)
    return const, b.emit_output(sv.gather(dyn), space, pos=dyn.astype(np.int64))


def fuse_spmv_jacobian(b, g_ins, jac_ins):
    """One pass over A for ``g = A @ phi(x)`` and ``J = A o phi'(x)`` (the two largest
    kernels read the same (A_ij, j) and gather phi_j / phi'_j at the same random j).

    Applies when the constraint values are ONE single-factor POLY whose slots are the even (value) slots of
    interleaved pair regions, the Jacobian values are ONE one-term-per-row POLY on the odd (derivative)
    slots, and every Jacobian entry is a term of g with the same coefficient, bit for bit.  The fused
    instruction replaces both in the union program only: a lone ``constraints`` / ``jacobian`` callback
    keeps its own kernel (no wasted Jacobian writes at rejected line-search points).
    Returns the new instruction or None."""
    tape = b.tape
    if len(g_ins) != 1 or len(jac_ins) != 1 or not b.pairs:
        return None
    P, Q = g_ins[0], jac_ins[0]
    if P.kind != T.K_POLY or Q.kind != T.K_POLY or P.accumulate or Q.accumulate:
        return None
    if np.any(P.f2 != NONE) or np.any(Q.f2 != NONE) or Q.coef.size != Q.count or P.coef.size < Builder.PAIR_MIN_NNZ:
        return None
    if not np.array_equal(Q.ptr, np.arange(Q.count + 1)):
        return None
    # slots must lie in pair regions: P on even (value) slots, Q on odd (derivative) slots
    bases = np.array(sorted(b.pairs), dtype=np.int64)
    ends = np.array([bs + 2 * b.pairs[bs][1] for bs in bases.tolist()], dtype=np.int64)

    def in_pairs(slots, parity):
        k = np.searchsorted(bases, slots, side="right") - 1
        ok = (k >= 0) & (slots < ends[np.maximum(k, 0)])
        return ok & (((slots - bases[np.maximum(k, 0)]) & 1) == parity)
    real = P.f1 != NONE
    if not np.all(in_pairs(P.f1[real], 0)) or not np.all(in_pairs(Q.f1, 1)):
        return None
    nsl = np.int64(tape.nslots + 2)
    prow = np.repeat(np.arange(P.count, dtype=np.int64), np.diff(P.ptr))
    if P.pos is not None:
        prow = np.asarray(P.pos, dtype=np.int64)[prow]
    qpos_of = np.arange(Q.count, dtype=np.int64) if Q.pos is None else np.asarray(Q.pos, dtype=np.int64)
    key_p = np.where(real, prow * nsl + P.f1, -1 - np.arange(P.coef.size, dtype=np.int64))
    key_q = tape.jac_rows[qpos_of].astype(np.int64) * nsl + (Q.f1 - 1)
    order = np.argsort(key_p, kind="stable")
    sk = key_p[order]
    if sk.size > 1 and np.any(sk[1:] == sk[:-1]):
        return None                                  # a repeated (row, column) in A: no one-to-one match
    at = np.searchsorted(sk, key_q)
    at = np.minimum(at, sk.size - 1)
    if not np.all(sk[at] == key_q):
        return None
    term = order[at]                                 # the term of g behind every Jacobian entry
    if not np.array_equal(P.coef[term].view(np.int64), Q.coef.view(np.int64)):
        return None
    qpos = np.full(P.coef.size, -1, dtype=np.int64)
    qpos[term] = qpos_of
    F = T.Instr(T.K_SPMVJ, dst_space=T.DST_G, dst_off=0, count=P.count, ptr=P.ptr, coef=P.coef, f1=P.f1,
                f2=P.f2, pos=P.pos, qpos=qpos.astype(np.int32), fused_jac=(P.id, Q.id))
    F.deps = tuple(sorted(set(P.deps) | set(Q.deps)))
    F.uses_lam = False
    F.dep_mask = P.dep_mask | Q.dep_mask
    F.level = max(P.level, Q.level)
    tape.add(F)
    return F


def _first_rejection_in_reference_order(prob, with_hessian):
    """The exception the REFERENCE raises first for a problem the rules reject.  Its oracle meets the rules in the
    order cyipopt asks: ``jacobianstructure`` (constraints in order, nlp_solver.py:309-335), ``hessianstructure``
    (objective, then constraints, :374-392), and only then values and the gradient; the compiler emits the value
    programs first.  A problem with two different defects (say ``quad_form`` of an affine expression - ValueError -
    and an atom without rules - NotImplementedError) must fail with the one the reference meets first.  Replayed on a
    scratch builder, only after the real compilation has failed."""
    b = Builder(prob)
    tape = b.tape
    steps = [lambda c=c: b.jac(c) for c in prob.constraints]
    if with_hessian:
        def hess_objective():
            b.in_objective_hessian = True
            try:
                b.hv(prob.objective, SymVec.slots([tape.sigma_slot]))
            finally:
                b.in_objective_hessian = False
        steps.append(hess_objective)
        coff = 0
        for con in prob.constraints:
            steps.append(lambda con=con, coff=coff: b.hv(con, SymVec.slot_range(tape.lam_slot + coff, con.size)))
            coff += con.size
    steps.append(lambda: b.value(prob.objective))
    steps.append(lambda: b.jac(prob.objective))
    steps += [lambda c=c: b.value(c) for c in prob.constraints]
    for step in steps:
        try:
            step()
        except Exception as e:          # noqa: BLE001
            return e
    return None


def compile_problem(prob, with_hessian=True):
    """ProblemIR -> Tape (host arrays only; ``GpuOracles`` uploads it through the C-ABI).

    ``with_hessian=False`` skips the Hessian program (the reference can serve first-order
    callbacks for expressions whose ``hess_vec`` rule rejects them; cyipopt then falls back to
    L-BFGS when the object has no usable ``hessian``).

    Rule-precondition failures raise what the reference raises (ValueError / NotImplementedError with its message,
    SURVEY 8b "Errors"), and when a problem has several, the one the reference's evaluation order meets first."""
    try:
        return _compile_problem(prob, with_hessian)
    except (ValueError, NotImplementedError, AttributeError, UnboundLocalError, TypeError, IndexError) as first:
        try:
            ref_first = _first_rejection_in_reference_order(prob, with_hessian)
        except Exception:               # noqa: BLE001
            ref_first = None
        if ref_first is None or (type(ref_first) is type(first) and str(ref_first) == str(first)):
            raise
        raise ref_first from first


def _compile_problem(prob, with_hessian=True):
    b = Builder(prob)
    tape = b.tape
    if tape.n_params:
        tape.param_values = prob.param_values()
    n, m = prob.n, prob.m
    var_ids = [v.attrs["id"] for v in prob.variables]
    off = b.var_off

    # ---- objective value ------------------------------------------------------
    fv = b.value(prob.objective)
    if fv.K != 1:
        raise ValueError("objective must be scalar")
    f_const, f_ins = _emit_output(b, fv, T.DST_F)
    tape.f_const = float(f_const[0])

    # ---- constraint values ------------------------------------------------------
    gv = SymVec.concat([b.value(c) for c in prob.constraints]) if prob.constraints else SymVec.zeros(0)
    tape.g_const, g_ins = _emit_output(b, gv, T.DST_G)

    # ---- gradient: grad_obj[offset + cols] = vals (assignment, last write wins) ----
    gd = b.jac(prob.objective)
    grad = SymVec.zeros(n)
    if gd:
        cols_all, parts = [], []
        for vid in var_ids:
            if vid in gd:
                _, c, v = gd[vid]
                cols_all.append(c + off[vid])
                parts.append(v)
        cols_all = np.concatenate(cols_all)
        vals = SymVec.concat(parts)
        last = np.full(n, -1, dtype=np.int64)
        last[cols_all] = np.arange(cols_all.size)
        dst = np.where(last >= 0)[0]
        grad = vals.gather(last[dst]).scatter_into(n, dst)
    tape.grad_const, grad_ins = _emit_output(b, grad, T.DST_GRAD)

    # ---- Jacobian triplets in emission order ----------------------------------
    R, C, V = [], [], []
    affine = [c.is_affine() for c in prob.constraints]
    coff = 0
    for con in prob.constraints:
        jd = b.jac(con)
        for vid in var_ids:
            if vid in jd:
                r, c, v = jd[vid]
                if not (np.size(r) == np.size(c) == v.K):
                    raise ValueError("row, column, and data array must all be the same length")
                R.append(r + coff)
                C.append(c + off[vid])
                V.append(v)
        coff += con.size
    R = np.concatenate(R) if R else np.zeros(0, np.int64)
    C = np.concatenate(C) if C else np.zeros(0, np.int64)
    jv = SymVec.concat(V)
    tape.jac_rows, tape.jac_cols = R.astype(np.int32), C.astype(np.int32)
    permutation_needed = not all(affine)
    tape.jac_is_list = not permutation_needed
    if permutation_needed and R.size:
        key = R * max(n, 1) + C
        srt = np.sort(key)                              # repeats are rare: look for them before paying for the inverse map
        if srt.size > 1 and bool(np.any(srt[1:] == srt[:-1])):     # quirk Q4: repeats receive the full sum
            uniq, inv = np.unique(key, return_inverse=True)
            jv = jv.group_sum(inv, uniq.size).gather(inv)
        del srt
    tape.jac_const, jac_ins = _emit_output(b, jv, T.DST_JAC)

    # ---- Hessian of the Lagrangian ----------------------------------------------
    HR, HC, HV = [], [], []

    def parse(hd):
        for v1 in var_ids:
            for v2 in var_ids:
                if (v1, v2) in hd:
                    r, c, v = hd[(v1, v2)]
                    if not (np.size(r) == np.size(c) == v.K):
                        # what scipy's coo_matrix raises inside sum_coo (nlp_solver.py:359-364) when a
                        # rule emits ragged triplets, e.g. multiply(promote(s), x) (binary_operators.py:529-535)
                        raise ValueError("row, column, and data array must all be the same length")
                    r = np.asarray(r, np.int64) + off[v1]
                    c = np.asarray(c, np.int64) + off[v2]
                    # sum_coo, then the lower triangle (nlp_solver.py:359-372).  Duplicates share their (row, col),
                    # so masking block by block BEFORE the sum gives the same entries in the same order and halves
                    # what a dense quad_form Hessian has to concatenate and sort.
                    low = np.flatnonzero(r >= c)
                    if low.size != r.size:
                        r, c, v = r[low], c[low], v.gather(low)
                    HR.append(r)
                    HC.append(c)
                    HV.append(v)

    if with_hessian:
        b.in_objective_hessian = True
        try:
            parse(b.hv(prob.objective, SymVec.slots([tape.sigma_slot])))
        finally:
            b.in_objective_hessian = False
        coff = 0
        for con in prob.constraints:
            parse(b.hv(con, SymVec.slot_range(tape.lam_slot + coff, con.size)))
            coff += con.size
    if HR:
        hr, hc, hv = Builder._coo_sum_duplicates(np.concatenate(HR), np.concatenate(HC), SymVec.concat(HV))
    else:
        hr = hc = np.zeros(0, np.int64)
        hv = SymVec.zeros(0)
    tape.hess_rows, tape.hess_cols = hr.astype(np.int32), hc.astype(np.int32)
    tape.hess_const, hess_ins = _emit_output(b, hv, T.DST_HESS)

    for name, ins in (("f", f_ins), ("grad", grad_ins), ("g", g_ins), ("jac", jac_ins), ("hess", hess_ins)):
        tape.programs[name] = _closure(tape, [i.id for i in ins])
    allp = set().union(*[set(p) for p in tape.programs.values()])
    fused = fuse_spmv_jacobian(b, g_ins, jac_ins)
    if fused is not None:
        allp = (allp - {g_ins[0].id, jac_ins[0].id}) | {fused.id}
    tape.programs["all"] = sorted(allp)
    return tape
