"""Dual recovery for ``prob.solve(nlp=True)``.

The reference returns no duals: ``IPOPT.invert`` builds ``Solution(status, opt_val, primal_vars, {}, attr)``
(cvxpy/reductions/solvers/nlp_solvers/ipopt_nlpif.py:100) although IPOPT hands back the constraint
multipliers (``info['mult_g']``) and every reduction above the solver already knows how to carry duals
home: ``Canonicalization.invert`` maps ``dual_vars`` through ``cons_id_map``
(cvxpy/reductions/canonicalization.py:76-84; Dnlp2Smooth is a Canonicalization), ``CvxAttr2Constr.invert``
does the same, ``FlipObjective.invert`` leaves them alone.  What is missing is the first link:

  * ``Bounds`` (nlp_solver.py:89-113) lowers every constraint of the smooth problem, IN ORDER, to
    ``g(x)`` with ``cl <= g <= cu``:  Equality lhs == rhs  ->  g = lhs - rhs, [0, 0];
    Inequality lhs <= rhs  ->  g = rhs - lhs, [0, inf);  NonPos(e)  ->  g = -e, [0, inf)
    (reductions/utilities.py:36-49); other constraint classes are dropped.
  * IPOPT's multipliers belong to the Lagrangian  f + mult_g' g.

So, per constraint of the smooth problem (ids known before the lowering):
    Equality     dual =  mult_g slice                 (f + nu' (lhs - rhs))
    Inequality   dual = -mult_g slice  (>= 0 at a minimiser: f + lam' (lhs - rhs))
    NonPos       dual = -mult_g slice
reshaped column-major to the constraint's shape.  ``record`` runs inside the (wrapped)
``NLPsolver._prepare_data_and_inv_data``, ``attach`` inside the (wrapped) solver ``invert``;
``dnlp_b200.nlp_solver.install`` installs both.
"""
import numpy as np


def record(smooth_problem, inverse_data):
    """Remember, on the solver stage's inverse data, which slice of g each smooth-problem constraint got."""
    table, off = [], 0
    for c in smooth_problem.constraints:
        kind = type(c).__name__
        if kind not in ("Equality", "Inequality", "NonPos"):
            continue                                 # Bounds drops every other class (nlp_solver.py:98-110)
        size = int(c.size)
        table.append((c.id, off, size, tuple(c.shape), 1.0 if kind == "Equality" else -1.0))
        off += size
    inverse_data.dnlp_dual_table = table
    inverse_data.dnlp_num_constraints = off
    return table


def duals_from_multipliers(inverse_data, mult_g):
    """{smooth-problem constraint id: dual value} from the solver's constraint multipliers."""
    table = getattr(inverse_data, "dnlp_dual_table", None)
    if table is None or mult_g is None:
        return {}
    lam = np.asarray(mult_g, dtype=np.float64).reshape(-1)
    if lam.size < inverse_data.dnlp_num_constraints:
        return {}
    out = {}
    for cid, off, size, shape, sign in table:
        v = sign * lam[off:off + size]
        out[cid] = v.reshape(shape, order="F") if shape else float(v[0])
    return out


def attach(solution_obj, raw_solution, inverse_data):
    """Fill ``dual_vars`` of the solver stage's Solution from ``raw_solution['mult_g']`` (IPOPT's info dict)."""
    mult = raw_solution.get("mult_g") if hasattr(raw_solution, "get") else None
    if mult is None or not getattr(solution_obj, "primal_vars", None):
        return solution_obj
    solution_obj.dual_vars = duals_from_multipliers(inverse_data, mult)
    return solution_obj
