"""Expression IR for smooth NLP problems.

This is the neutral, serialisable form of what the reference hands to its
``Oracles`` object: the *smooth* problem produced by ``Dnlp2Smooth`` and
``Bounds`` (reference: cvxpy/reductions/solvers/nlp_solvers/nlp_solver.py:61-79,
81-178).  A node is one atom of the reference's expression tree; op names follow
the reference's class names so that every derivative rule can cite the reference
file it mirrors.

Two producers exist:
  * ``dnlp_b200.frontend_cvxpy.problem_to_ir`` walks a live reference problem;
  * the small builder API below (operator overloading on ``Node``) lets tests
    and ``bench.py`` construct the same trees without the reference installed.

Layout conventions (reference: nlp_solver.py:205-210): every expression value is
flattened in column-major ('F') order; variables are laid out in
``Problem.variables()`` order (objective first, then constraints, first
appearance; cvxpy/problems/problem.py:410-421).
"""
from __future__ import annotations

import itertools
import json
from fractions import Fraction

import numpy as np
import scipy.sparse as sp

ELEMENTWISE_UNARY = (
    "exp", "log", "entr", "logistic", "power", "sin", "cos", "tan",
    "sinh", "tanh", "asinh", "atanh", "xexp",
)
AFFINE_OPS = (
    "add", "neg", "sum", "index", "special_index", "reshape", "transpose",
    "promote", "broadcast_to",
)
BILINEAR_OPS = ("multiply", "matmul")
OTHER_SMOOTH = ("quad_form", "quad_over_lin", "rel_entr")
# "unsupported": an atom of the reference without NLP rules, kept so that the compiler raises where the reference would
ALL_OPS = ("var", "const", "param", "unsupported") + AFFINE_OPS + BILINEAR_OPS + ELEMENTWISE_UNARY + OTHER_SMOOTH

_var_ids = itertools.count(1)


def _size(shape):
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


class Node:
    """One atom.  ``shape`` is the cvxpy shape tuple, ``attrs`` the atom data."""

    __array_priority__ = 1000  # numpy defers to our __r*__ operators
    __slots__ = ("op", "args", "shape", "attrs", "_vars", "_params")

    def __init__(self, op, args=(), shape=(), **attrs):
        if op not in ALL_OPS:
            raise NotImplementedError(
                "Atom %s does not have a Jacobian, or it has not been implemented yet." % op)
        self.op = op
        self.args = list(args)
        self.shape = tuple(int(s) for s in shape)
        self.attrs = attrs
        self._vars = None
        self._params = None

    # ---- structural predicates (mirror Expression.is_constant / is_affine) ----
    @property
    def size(self):
        return _size(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    def is_scalar(self):
        return self.size == 1

    def variables(self):
        """First-appearance ordered unique variables (utilities/canonical.py:58-62)."""
        if self._vars is None:
            if self.op == "var":
                self._vars = [self]
            else:
                seen, out = set(), []
                for a in self.args:
                    for v in a.variables():
                        if id(v) not in seen:
                            seen.add(id(v))
                            out.append(v)
                self._vars = out
        return self._vars

    def is_constant(self):
        return self.op != "var" and (0 in self.shape or not self.variables())

    def params(self):
        """First-appearance ordered unique ``param`` leaves (cvxpy Parameters: constants whose VALUE may change
        between solves, expressions/constants/parameter.py:35)."""
        if self._params is None:
            if self.op == "param":
                self._params = [self]
            else:
                seen, out = set(), []
                for a in self.args:
                    for q in a.params():
                        if id(q) not in seen:
                            seen.add(id(q))
                            out.append(q)
                self._params = out
        return self._params

    def has_params(self):
        return bool(self.params())

    def is_var(self):
        return self.op == "var"

    def is_affine(self):
        """Structural twin of ``Expression.is_affine`` for the atoms that reach the oracle."""
        if self.op == "var" or self.is_constant():
            return True
        if self.op in AFFINE_OPS:
            return all(a.is_affine() for a in self.args)
        if self.op in BILINEAR_OPS:
            a, b = self.args
            return (a.is_constant() and b.is_affine()) or (b.is_constant() and a.is_affine())
        if self.op == "power":
            # power(x, 1) is DCP-affine; power_canon removes it before the oracle sees it
            return self.attrs["p"] == 1 and self.args[0].is_affine()
        if self.op == "unsupported":
            # an atom of the reference without NLP rules (frontend_cvxpy.py): it only exists so that the compiler can
            # raise NotImplementedError WHERE the reference would (rules.Builder.jac / hv), not at conversion time
            return bool(self.attrs.get("affine", False))
        return False

    @property
    def value(self):
        """Constant payload (dense ndarray or scipy sparse)."""
        return self.attrs["value"]

    # ---- builder API (matches Expression.__add__/__mul__/... broadcasting) ----
    def __add__(self, other):
        return add(self, other)

    def __radd__(self, other):
        return add(other, self)

    def __sub__(self, other):
        return add(self, neg(as_node(other)))

    def __rsub__(self, other):
        return add(other, neg(self))

    def __neg__(self):
        return neg(self)

    def __mul__(self, other):
        return _star(self, as_node(other))

    def __rmul__(self, other):
        return _star(as_node(other), self)

    def __matmul__(self, other):
        return matmul(self, other)

    def __rmatmul__(self, other):
        return matmul(other, self)

    def __pow__(self, p):
        return power(self, p)

    def __getitem__(self, key):
        return index(self, key)

    @property
    def T(self):
        return transpose(self)

    def __repr__(self):
        if self.op == "var":
            return "var%d%s" % (self.attrs["id"], list(self.shape))
        if self.op == "const":
            return "const%s" % (list(self.shape),)
        if self.op == "param":
            return "param%d%s" % (self.attrs["id"], list(self.shape))
        return "%s(%s)" % (self.op, ", ".join(repr(a) for a in self.args))


# --------------------------------------------------------------------------
# leaves
# --------------------------------------------------------------------------
def Variable(shape=(), name=None, lb=None, ub=None, value=None):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    vid = next(_var_ids)
    return Node("var", (), shape, id=vid, name=name or "var%d" % vid, lb=lb, ub=ub, value=value)


def Constant(value):
    if sp.issparse(value):
        v = sp.csc_array(value).astype(np.float64)
        return Node("const", (), v.shape, value=v)
    v = np.asarray(value, dtype=np.float64)
    return Node("const", (), v.shape, value=v)


def Parameter(shape=(), value=None, name=None):
    """A constant whose value can change between solves WITHOUT recompiling the tape: it lives in the value
    buffer next to x / sigma / lambda (``GpuOracles.set_parameters``)."""
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    pid = next(_var_ids)
    v = np.zeros(shape) if value is None else np.asarray(value, dtype=np.float64).reshape(shape)
    return Node("param", (), shape, id=pid, name=name or "param%d" % pid, value=v)


def as_node(x):
    return x if isinstance(x, Node) else Constant(x)


def eval_constant(node):
    """Numeric value of a constant subtree (constants and parameters at their current values)."""
    if node.op in ("const", "param"):
        return node.attrs["value"]
    from . import _consteval
    return _consteval.numeric(node, [eval_constant(a) for a in node.args])


# --------------------------------------------------------------------------
# affine atoms
# --------------------------------------------------------------------------
def _broadcast2(a, b):
    """Expression.broadcast (cvxpy/expressions/expression.py:681-713)."""
    a, b = as_node(a), as_node(b)
    if a.is_scalar() and not b.is_scalar():
        a = promote(a, b.shape)
    elif b.is_scalar() and not a.is_scalar():
        b = promote(b, a.shape)
    elif a.is_scalar() and b.is_scalar():
        return a, b
    if a.ndim == 2 and b.ndim == 2:
        dims = [max(a.shape[i], b.shape[i]) for i in range(2)]
        if a.shape[0] == 1 and a.shape[0] < dims[0]:
            a = matmul(np.ones((dims[0], 1)), a)
        if b.shape[0] == 1 and b.shape[0] < dims[0]:
            b = matmul(np.ones((dims[0], 1)), b)
        if a.shape[1] == 1 and a.shape[1] < dims[1]:
            a = matmul(a, np.ones((1, dims[1])))
        if b.shape[1] == 1 and b.shape[1] < dims[1]:
            b = matmul(b, np.ones((1, dims[1])))
    elif a.ndim != b.ndim:
        out = np.broadcast_shapes(a.shape, b.shape)
        if a.shape != out:
            a = broadcast_to(a, out)
        if b.shape != out:
            b = broadcast_to(b, out)
    return a, b


def _const_fold(op, args, shape, **attrs):
    """Build a node; if every arg is constant evaluate it right away."""
    node = Node(op, args, shape, **attrs)
    if all(a.op == "const" for a in node.args) and node.args:
        from . import _consteval
        return Constant(_consteval.numeric(node, [a.attrs["value"] for a in node.args]))
    return node


def add(*terms):
    flat = []
    terms = [as_node(t) for t in terms]
    if len(terms) == 2:
        terms = list(_broadcast2(*terms))
    for t in terms:
        flat.extend(t.args if t.op == "add" else [t])
    shape = np.broadcast_shapes(*[t.shape for t in flat])
    return _const_fold("add", flat, shape)


def neg(x):
    x = as_node(x)
    return _const_fold("neg", [x], x.shape)


def sum(x, axis=None, keepdims=False):  # noqa: A001  (mirrors cvxpy.sum)
    x = as_node(x)
    shape = np.sum(np.zeros(x.shape), axis=axis, keepdims=keepdims).shape
    return _const_fold("sum", [x], shape, axis=axis, keepdims=keepdims)


def _format_slice(k, dim):
    if isinstance(k, (int, np.integer)):
        k = int(k)
        if k < 0:
            k += dim
        if not 0 <= k < dim:
            raise IndexError("Index out of bounds.")
        return (k, k + 1, 1)
    start, stop, step = k.indices(dim)
    if step < 0 and stop < 0:
        stop = None
    return (start, stop, step)


def index(x, key):
    x = as_node(x)
    orig = key
    tkey = key if isinstance(key, tuple) else (key,)
    if any(isinstance(k, (list, np.ndarray)) for k in tkey):
        return special_index(x, key)
    if len(tkey) < x.ndim:
        tkey = tkey + (slice(None),) * (x.ndim - len(tkey))
    if len(tkey) > x.ndim:
        raise IndexError("Too many indices for expression.")
    fkey = [_format_slice(k, d) for k, d in zip(tkey, x.shape)]
    shape = np.empty(x.shape, dtype=np.dtype([]))[orig].shape
    return _const_fold("index", [x], shape, key=fkey, orig_key=_encode_key(orig))


def special_index(x, key):
    x = as_node(x)
    idx_mat = np.arange(x.size).reshape(x.shape, order="F")
    sel = idx_mat[key]
    return _const_fold("special_index", [x], sel.shape, select=np.asarray(sel, dtype=np.int64))


def reshape(x, shape, order="F"):
    x = as_node(x)
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    shape = tuple(shape)
    if -1 in shape:
        shape = np.empty(x.shape).reshape(shape).shape
    return _const_fold("reshape", [x], shape, order=order)


def vec(x):
    return reshape(x, (as_node(x).size,), "F")


def transpose(x, axes=None):
    x = as_node(x)
    shape = np.transpose(np.empty(x.shape), axes).shape
    return _const_fold("transpose", [x], shape, axes=axes)


def promote(x, shape):
    x = as_node(x)
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    if x.shape == tuple(shape):
        return x
    return _const_fold("promote", [x], tuple(shape))


def broadcast_to(x, shape):
    x = as_node(x)
    return _const_fold("broadcast_to", [x], tuple(shape))


# --------------------------------------------------------------------------
# products
# --------------------------------------------------------------------------
def multiply(a, b):
    a, b = _broadcast2(a, b)
    return _const_fold("multiply", [a, b], np.broadcast_shapes(a.shape, b.shape))


def _mul_shape(a, b):
    if a.ndim == 0 or b.ndim == 0:
        return b.shape if a.ndim == 0 else a.shape
    return np.matmul(np.zeros(a.shape), np.zeros(b.shape)).shape


def matmul(a, b):
    a, b = as_node(a), as_node(b)
    if a.ndim > 2 or b.ndim > 2:
        raise ValueError("Multiplication with N-d arrays is not yet supported")
    return _const_fold("matmul", [a, b], _mul_shape(a, b))


def _star(a, b):
    """Expression.__mul__ (expression.py:745-776): scalars scale, otherwise matmul."""
    if a.shape == () or b.shape == ():
        return multiply(a, b)
    if a.shape[-1] != b.shape[0] and (a.is_scalar() or b.is_scalar()):
        return multiply(a, b)
    return matmul(a, b)


# --------------------------------------------------------------------------
# smooth atoms
# --------------------------------------------------------------------------
def _elementwise(op):
    def build(x):
        x = as_node(x)
        return _const_fold(op, [x], x.shape)
    build.__name__ = op
    return build


exp = _elementwise("exp")
log = _elementwise("log")
entr = _elementwise("entr")
logistic = _elementwise("logistic")
sin = _elementwise("sin")
cos = _elementwise("cos")
tan = _elementwise("tan")
sinh = _elementwise("sinh")
tanh = _elementwise("tanh")
asinh = _elementwise("asinh")
atanh = _elementwise("atanh")
xexp = _elementwise("xexp")


def rational_power(p, max_denom=1024):
    """``p_rational`` as the reference derives it (atoms/elementwise/power.py:150-182,
    utilities/power_tools.py:106-149): integers stay ints, p>1 and p<0 and 0<p<1
    become ``Fraction(p).limit_denominator(1024)``."""
    if isinstance(p, (int, np.integer)) or float(p).is_integer():
        return int(p)
    return Fraction(p).limit_denominator(max_denom)


def power(x, p, p_rational="auto"):
    x = as_node(x)
    if p_rational == "auto":
        p_rational = rational_power(p)
    return _const_fold("power", [x], x.shape, p=float(p), p_rational=p_rational)


def square(x):
    return power(x, 2)


def rel_entr(x, y):
    x, y = as_node(x), as_node(y)
    return _const_fold("rel_entr", [x, y], np.broadcast_shapes(x.shape, y.shape))


def quad_over_lin(x, y):
    return _const_fold("quad_over_lin", [as_node(x), as_node(y)], ())


def quad_form(x, P):
    x, P = as_node(x), as_node(P)
    return _const_fold("quad_form", [x, P], ())


# --------------------------------------------------------------------------
# problem container
# --------------------------------------------------------------------------
class ProblemIR:
    """Smooth NLP in the form ``Oracles`` sees it.

    minimize objective(x)  s.t.  cl <= [c.flatten('F') for c in constraints] <= cu,
    lb <= x <= ub.  ``variables`` is the flat layout order.
    """

    def __init__(self, objective, constraints=(), variables=None,
                 cl=None, cu=None, lb=None, ub=None, x0=None):
        self.objective = as_node(objective)
        self.constraints = [as_node(c) for c in constraints]
        if variables is None:
            seen, variables = set(), []
            for e in [self.objective] + self.constraints:
                for v in e.variables():
                    if id(v) not in seen:
                        seen.add(id(v))
                        variables.append(v)
        self.variables = list(variables)
        self.n = int(np.sum([v.size for v in self.variables], dtype=np.int64)) if self.variables else 0
        self.m = int(np.sum([c.size for c in self.constraints], dtype=np.int64)) if self.constraints else 0
        seen, self.params = set(), []
        for e in [self.objective] + self.constraints:
            for q in e.params():
                if id(q) not in seen:
                    seen.add(id(q))
                    self.params.append(q)
        self.n_params = int(np.sum([q.size for q in self.params], dtype=np.int64)) if self.params else 0
        self.cl = None if cl is None else np.asarray(cl, dtype=np.float64)
        self.cu = None if cu is None else np.asarray(cu, dtype=np.float64)
        self.lb = None if lb is None else np.asarray(lb, dtype=np.float64)
        self.ub = None if ub is None else np.asarray(ub, dtype=np.float64)
        self.x0 = None if x0 is None else np.asarray(x0, dtype=np.float64)

    def param_values(self):
        """Current parameter values, flattened column-major in ``self.params`` order."""
        if not self.params:
            return np.zeros(0)
        return np.concatenate([np.asarray(q.attrs["value"], dtype=np.float64).flatten(order="F") for q in self.params])

    def folded(self):
        """The same problem with every parameter-dependent constant subtree evaluated at the current values
        (what the reference's rules see through ``.value``): input of the CPU oracle."""
        memo = {}

        def visit(n):
            if id(n) in memo:
                return memo[id(n)]
            if n.op == "var" or not n.has_params():
                out = n
            elif n.is_constant():
                v = np.asarray(eval_constant(n), dtype=np.float64)
                out = Constant(v.reshape(n.shape, order="F") if v.size == n.size else v)
            else:
                out = Node(n.op, [visit(a) for a in n.args], n.shape, **n.attrs)
            memo[id(n)] = out
            return out
        return ProblemIR(visit(self.objective), [visit(c) for c in self.constraints], self.variables,
                         self.cl, self.cu, self.lb, self.ub, self.x0)

    def var_offsets(self):
        off, out = 0, {}
        for v in self.variables:
            out[id(v)] = off
            off += v.size
        return out


# --------------------------------------------------------------------------
# (de)serialisation -- golden fixtures travel as one .npz per problem
# --------------------------------------------------------------------------
def _encode_key(key):
    def enc(k):
        if isinstance(k, slice):
            return {"s": [None if v is None else int(v) for v in (k.start, k.stop, k.step)]}
        if k is None:
            return {"n": 1}
        if k is Ellipsis:
            return {"e": 1}
        return {"i": int(k)}
    if isinstance(key, tuple):
        return {"t": [enc(k) for k in key]}
    return enc(key)


def decode_key(d):
    def dec(k):
        if "s" in k:
            return slice(*k["s"])
        if "n" in k:
            return None
        if "e" in k:
            return Ellipsis
        return k["i"]
    if "t" in d:
        return tuple(dec(k) for k in d["t"])
    return dec(d)


def dump_problem(prob):
    """-> (json_str, {array_name: ndarray}) ; DAG sharing of nodes is preserved."""
    arrays, nodes, memo = {}, [], {}

    def put(arr):
        name = "a%d" % len(arrays)
        arrays[name] = arr
        return name

    def visit(n):
        if id(n) in memo:
            return memo[id(n)]
        kids = [visit(a) for a in n.args]
        attrs = {}
        for k, v in n.attrs.items():
            if k == "value" and n.op == "const":
                if sp.issparse(v):
                    c = sp.coo_array(v)
                    attrs[k] = {"coo": [put(np.asarray(c.coords[0], np.int64)),
                                        put(np.asarray(c.coords[1], np.int64)),
                                        put(np.asarray(c.data, np.float64))],
                                "shape": list(v.shape)}
                else:
                    attrs[k] = {"dense": put(np.asarray(v, np.float64))}
            elif isinstance(v, np.ndarray):
                attrs[k] = {"nd": put(v)}
            elif isinstance(v, Fraction):
                attrs[k] = {"frac": [v.numerator, v.denominator]}
            elif isinstance(v, tuple):
                attrs[k] = {"tuple": list(v)}
            else:
                attrs[k] = v
        nodes.append({"op": n.op, "args": kids, "shape": list(n.shape), "attrs": attrs})
        memo[id(n)] = len(nodes) - 1
        return memo[id(n)]

    doc = {
        "objective": visit(prob.objective),
        "constraints": [visit(c) for c in prob.constraints],
        "variables": [visit(v) for v in prob.variables],
    }
    doc["nodes"] = nodes
    for k in ("cl", "cu", "lb", "ub", "x0"):
        v = getattr(prob, k)
        doc[k] = None if v is None else put(v)
    return json.dumps(doc), arrays


def load_problem(json_str, arrays):
    doc = json.loads(json_str)
    built = []
    for nd in doc["nodes"]:
        attrs = {}
        for k, v in nd["attrs"].items():
            if isinstance(v, dict) and "dense" in v:
                attrs[k] = np.asarray(arrays[v["dense"]], np.float64)
            elif isinstance(v, dict) and "coo" in v:
                r, c, d = (arrays[a] for a in v["coo"])
                attrs[k] = sp.csc_array(sp.coo_array((d, (r, c)), shape=tuple(v["shape"])))
            elif isinstance(v, dict) and "nd" in v:
                attrs[k] = arrays[v["nd"]]
            elif isinstance(v, dict) and "frac" in v:
                attrs[k] = Fraction(*v["frac"])
            elif isinstance(v, dict) and "tuple" in v:
                attrs[k] = tuple(v["tuple"])
            else:
                attrs[k] = v
        built.append(Node(nd["op"], [built[i] for i in nd["args"]], nd["shape"], **attrs))
    get = lambda k: None if doc[k] is None else arrays[doc[k]]  # noqa: E731
    return ProblemIR(built[doc["objective"]], [built[i] for i in doc["constraints"]],
                     [built[i] for i in doc["variables"]],
                     get("cl"), get("cu"), get("lb"), get("ub"), get("x0"))
