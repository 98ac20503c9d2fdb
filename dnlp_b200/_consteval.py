"""NumPy evaluation of *constant* subtrees, used once at build/compile time.

The reference keeps constant trees in the problem and re-evaluates them with
NumPy on every callback (cvxpy/atoms/atom.py:431-449); folding them once at
compile time yields the same numbers.  This is host-side compile-time code, not
an evaluation fallback: nothing that depends on a variable ever goes through it.
"""
import numpy as np
import scipy.sparse as sp
from scipy.special import rel_entr as _rel_entr, xlogy as _xlogy


def _dense(v):
    return v.toarray() if sp.issparse(v) else np.asarray(v, dtype=np.float64)


def numeric(node, values):
    op = node.op
    if op == "add":
        out = values[0]
        for v in values[1:]:
            out = out + v
        return _dense(out)
    if op == "neg":
        return -values[0]
    if op == "sum":
        v = _dense(values[0])
        return np.sum(v, axis=node.attrs["axis"], keepdims=node.attrs["keepdims"])
    if op == "index":
        from .ir import decode_key
        return _dense(values[0])[decode_key(node.attrs["orig_key"])]
    if op == "special_index":
        return _dense(values[0]).flatten(order="F")[node.attrs["select"]]
    if op == "reshape":
        return np.reshape(_dense(values[0]), node.shape, order=node.attrs["order"])
    if op == "transpose":
        return np.transpose(_dense(values[0]), node.attrs["axes"])
    if op == "promote":
        return np.ones(node.shape) * _dense(values[0])
    if op == "broadcast_to":
        return np.broadcast_to(_dense(values[0]), node.shape).copy()
    if op == "multiply":
        if sp.issparse(values[0]):
            return values[0].multiply(values[1])
        if sp.issparse(values[1]):
            return values[1].multiply(values[0])
        return np.multiply(values[0], values[1])
    if op == "matmul":
        a, b = values
        if np.ndim(a) == 0 or np.ndim(b) == 0:
            return a * b
        return a @ b
    x = _dense(values[0])
    if op == "exp":
        return np.exp(x)
    if op == "log":
        return np.log(x)
    if op == "entr":
        r = np.asarray(-_xlogy(x, x), dtype=np.float64)
        r[np.isnan(r)] = -np.inf
        return r
    if op == "logistic":
        return np.logaddexp(0, x)
    if op == "power":
        return np.power(x, node.attrs["p"])
    if op in ("sin", "cos", "tan", "sinh", "tanh"):
        return getattr(np, op)(x)
    if op == "asinh":
        return np.arcsinh(x)
    if op == "atanh":
        return np.arctanh(x)
    if op == "xexp":
        return x * np.exp(x)
    if op == "rel_entr":
        return _rel_entr(x, _dense(values[1]))
    if op == "quad_over_lin":
        return np.square(x).sum() / _dense(values[1])
    if op == "quad_form":
        return np.dot(x.T, values[1].dot(x))
    raise NotImplementedError(op)
