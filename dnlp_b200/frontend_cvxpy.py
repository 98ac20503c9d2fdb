"""Reference problem -> IR.

Walks the smooth problem the reference builds in ``NLPsolver.apply``
(cvxpy/reductions/solvers/nlp_solvers/nlp_solver.py:61-79) and emits
``dnlp_b200.ir`` nodes.  Dispatch is on class *names*, so this module never
imports cvxpy: it works with whichever copy of the reference the caller has
loaded, and the rest of the package stays importable without it.
"""
import numpy as np
import scipy.sparse as sp

from . import ir

_ELEMENTWISE = {name: name for name in ir.ELEMENTWISE_UNARY if name != "power"}
_SIMPLE = {
    "AddExpression": "add", "NegExpression": "neg", "Promote": "promote",
    "multiply": "multiply", "MulExpression": "matmul", "QuadForm": "quad_form",
    "quad_over_lin": "quad_over_lin", "rel_entr": "rel_entr",
}
_SIMPLE.update(_ELEMENTWISE)


def _const(expr):
    v = expr.value
    if v is None:
        raise ValueError("constant/parameter without a value: %s" % expr)
    if sp.issparse(v):
        node = ir.Constant(v)
    else:
        node = ir.Constant(np.asarray(v, dtype=np.float64))
    # the reference's shape wins (e.g. a (n,) constant stored as a column)
    if node.shape != tuple(expr.shape) and node.size == int(np.prod(expr.shape, dtype=np.int64)):
        node = ir.Node("const", (), expr.shape,
                       value=np.reshape(node.attrs["value"] if not sp.issparse(node.attrs["value"])
                                        else node.attrs["value"].toarray(), expr.shape, order="F"))
    return node


# Atoms under which a Parameter may stay a value slot: sums, elementwise products and the affine index /
# shape atoms (dnlp_b200/rules.py handles their value, Jacobian and Hessian rules with a symbolic constant
# factor).  Anywhere else (the matrix of a product with variables, quad_form's matrix, an exponent) the
# parameter is FROZEN at its current value: correct, and a new value then means a new tape (the fingerprint
# of a frozen parameter includes its value).
_PARAM_OK_PARENTS = ("AddExpression", "NegExpression", "multiply", "Promote", "Sum", "index", "special_index",
                     "reshape", "transpose", "broadcast_to")


def _param_tree(expr, memo):
    """IR of a constant sub-expression that contains Parameters, or None when some atom in it cannot carry a
    parameter slot (then the caller folds the whole sub-expression to its current value)."""
    cls = type(expr).__name__
    if cls == "Parameter":
        key = ("param", expr.id)
        if key not in memo:
            if expr.value is None:
                raise ValueError("constant/parameter without a value: %s" % expr)
            memo[key] = ir.Node("param", (), expr.shape, id=int(expr.id), name=expr.name(),
                                value=np.asarray(expr.value, dtype=np.float64))
        return memo[key]
    if not expr.parameters():
        return _const(expr)
    if cls not in _PARAM_OK_PARENTS and cls != "MulExpression":
        return None
    kids = [_param_tree(a, memo) for a in expr.args]
    if any(k is None for k in kids):
        return None
    if cls == "MulExpression":                    # constant @ parameter-tree (an affine map of the parameters)
        if expr.args[0].parameters() or not expr.args[1].parameters():
            return None
        return ir.Node("matmul", kids, expr.shape)
    if cls == "Sum":
        return ir.Node("sum", kids, expr.shape, axis=expr.axis, keepdims=bool(expr.keepdims))
    if cls == "index":
        fkey = [(int(s.start), None if s.stop is None else int(s.stop), int(s.step)) for s in expr.key]
        return ir.Node("index", kids, expr.shape, key=fkey, orig_key=ir._encode_key(expr._orig_key))
    if cls == "special_index":
        return ir.Node("special_index", kids, expr.shape, select=np.asarray(expr._select_mat, dtype=np.int64))
    if cls == "reshape":
        return ir.Node("reshape", kids, expr.shape, order=expr.order)
    if cls == "transpose":
        axes = getattr(expr, "axes", None)
        return ir.Node("transpose", kids, expr.shape, axes=None if axes is None else tuple(int(a) for a in axes))
    return ir.Node(_SIMPLE.get(cls, {"Promote": "promote", "broadcast_to": "broadcast_to"}.get(cls)), kids, expr.shape)


def expr_to_ir(expr, memo, parent=None):
    key = id(expr)
    if key in memo:
        return memo[key]
    cls = type(expr).__name__
    if expr.is_constant() and cls != "Variable" and expr.parameters() and (parent is None or parent in _PARAM_OK_PARENTS):
        node = _param_tree(expr, memo)
        if node is not None:
            memo[key] = node
            return node
    if cls == "Variable":
        bounds = getattr(expr, "bounds", None)
        node = ir.Node("var", (), expr.shape, id=int(expr.id), name=expr.name(),
                       lb=None if not bounds else bounds[0], ub=None if not bounds else bounds[1],
                       value=None if expr.value is None else np.asarray(expr.value, np.float64))
    elif expr.is_constant():
        node = _const(expr)
    else:
        args = [expr_to_ir(a, memo, parent=cls) for a in expr.args]
        if cls in _SIMPLE:
            node = ir.Node(_SIMPLE[cls], args, expr.shape)
        elif cls == "Sum":
            node = ir.Node("sum", args, expr.shape, axis=expr.axis, keepdims=bool(expr.keepdims))
        elif cls == "index":
            fkey = [(int(s.start), None if s.stop is None else int(s.stop), int(s.step))
                    for s in expr.key]
            node = ir.Node("index", args, expr.shape, key=fkey,
                           orig_key=ir._encode_key(expr._orig_key))
        elif cls == "special_index":
            node = ir.Node("special_index", args, expr.shape,
                           select=np.asarray(expr._select_mat, dtype=np.int64))
        elif cls == "reshape":
            node = ir.Node("reshape", args, expr.shape, order=expr.order)
        elif cls in ("transpose",):
            axes = getattr(expr, "axes", None)
            node = ir.Node("transpose", args, expr.shape,
                           axes=None if axes is None else tuple(int(a) for a in axes))
        elif cls == "broadcast_to":
            node = ir.Node("broadcast_to", args, expr.shape)
        elif cls == "power":
            p = expr.p.value
            if p is None and expr.p_rational is None:
                raise ValueError("Argument error in jacobian for atom power.")
            node = ir.Node("power", args, expr.shape,
                           p=float(p), p_rational=expr.p_rational)
        else:
            # no NLP rules in the reference (atoms/atom.py:587-593).  The reference only finds out when a structure
            # pass reaches the atom, AFTER whatever it meets earlier in its own evaluation order: keep the node and
            # let the compiler raise there (compiler.compile_problem replays that order when a problem is rejected)
            node = ir.Node("unsupported", args, expr.shape, cls=cls, affine=bool(expr.is_affine()))
        if node.is_affine() != bool(expr.is_affine()):
            raise AssertionError("affine classification differs from the reference for %s" % cls)
    memo[key] = node
    return node


def problem_to_ir(problem, cl=None, cu=None, lb=None, ub=None, x0=None):
    """``problem`` is ``Bounds.new_problem`` (constraints already Zero/NonNeg)."""
    memo = {}
    obj = expr_to_ir(problem.objective.args[0], memo)
    cons = [expr_to_ir(c.args[0], memo) for c in problem.constraints]
    variables = [expr_to_ir(v, memo) for v in problem.variables()]
    return ir.ProblemIR(obj, cons, variables, cl, cu, lb, ub, x0)


def data_to_ir(data):
    """From the dict returned by the reference's ``NLPsolver.apply`` chain."""
    return problem_to_ir(data["problem"], data["cl"], data["cu"], data["lb"], data["ub"], data["x0"])
