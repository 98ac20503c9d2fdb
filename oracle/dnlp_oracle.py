"""CPU restatement of the reference NLP oracle (TEST INFRASTRUCTURE ONLY).

This module restates, rule by rule, what cvxgrp/DNLP's ``Oracles`` object and the
per-atom ``numeric`` / ``_jacobian`` / ``_hess_vec`` rules compute, on top of the
neutral IR in ``dnlp_b200.ir``.  It is the *checker* for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it.  Nothing under ``dnlp_b200/`` imports it, and it is never a fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the live reference
from /root/reference, runs its seven callbacks on a suite of problems and stores
IR + inputs + outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this restatement against every one of them (structures bit-exact, values
rel 1e-10).

Every function cites the reference lines it follows (paths relative to
/root/reference/cvxpy).  Third-party arithmetic on the path (NumPy ufuncs, SciPy
sparse products / ``sum_duplicates`` / ``scipy.special``) is called directly, as
the reference does, so triplet *order* is pinned by running the same library
calls.
"""
import numpy as np
import scipy.sparse as sp
from scipy.special import rel_entr as rel_entr_scipy
from scipy.special import xlogy

from dnlp_b200 import ir as _ir


def _flatF(v):
    return np.asarray(v).flatten(order="F")


# ---------------------------------------------------------------------------
# values: Atom._value_impl / numeric  (atoms/atom.py:431-449)
# ---------------------------------------------------------------------------
def numeric(node, env):
    """Value of ``node``; ``env`` maps variable id -> ndarray of the variable's shape."""
    op = node.op
    if op == "var":
        return env[node.attrs["id"]]
    if op == "const":
        return node.attrs["value"]
    v = [numeric(a, env) for a in node.args]
    if op == "add":                       # affine/add_expr.py:72-73
        out = v[0]
        for t in v[1:]:
            out = out + t
        return out
    if op == "neg":                       # affine/unary_operators.py:37
        return -v[0]
    if op == "sum":                       # affine/sum.py:93-101
        if sp.issparse(v[0]):
            r = np.asarray(v[0].sum(axis=node.attrs["axis"]))
            if not node.attrs["keepdims"] and node.attrs["axis"] is not None:
                r = r.flatten()
            return r
        return np.sum(v[0], axis=node.attrs["axis"], keepdims=node.attrs["keepdims"])
    if op == "index":                     # affine/index.py:88-90
        return v[0][_ir.decode_key(node.attrs["orig_key"])]
    if op == "special_index":             # affine/index.py:194-197 (via the select matrix)
        return _flatF(v[0])[node.attrs["select"]]
    if op == "reshape":                   # affine/reshape.py:102
        return np.reshape(v[0], node.shape, order=node.attrs["order"])
    if op == "transpose":                 # affine/transpose.py:53
        return np.transpose(v[0], node.attrs["axes"])
    if op == "promote":                   # affine/promote.py:68-70
        return np.ones(node.shape) * v[0]
    if op == "broadcast_to":              # affine/broadcast_to.py:45-46
        return np.broadcast_to(v[0], node.shape)
    if op == "multiply":                  # affine/binary_operators.py:431-438
        if sp.issparse(v[0]):
            return v[0].multiply(v[1])
        if sp.issparse(v[1]):
            return v[1].multiply(v[0])
        return np.multiply(v[0], v[1])
    if op == "matmul":                    # affine/binary_operators.py:134-140
        if np.shape(v[0]) == () or np.shape(v[1]) == ():
            return v[0] * v[1]
        return v[0] @ v[1]
    x = v[0]
    if op == "exp":                       # elementwise/exp.py:34-35
        return np.exp(x)
    if op == "log":                       # elementwise/log.py:33-36
        return np.log(x)
    if op == "entr":                      # elementwise/entr.py:35-44
        r = -xlogy(x, x)
        if np.isscalar(r):
            return -np.inf if np.isnan(r) else r
        r = np.asarray(r)
        r[np.isnan(r)] = -np.inf
        return r
    if op == "logistic":                  # elementwise/logistic.py:36-39
        return np.logaddexp(0, x)
    if op == "power":                     # elementwise/power.py:187-188 (exact p, quirk Q3)
        return np.power(x, node.attrs["p"])
    if op == "sin":                       # elementwise/trig.py:33-36
        return np.sin(x)
    if op == "cos":                       # elementwise/trig.py:113-116
        return np.cos(x)
    if op == "tan":                       # elementwise/trig.py:194-197
        return np.tan(x)
    if op == "sinh":                      # elementwise/hyperbolic.py:33-36
        return np.sinh(x)
    if op == "tanh":                      # elementwise/hyperbolic.py:108-111
        return np.tanh(x)
    if op == "asinh":                     # elementwise/hyperbolic.py:183-186
        return np.arcsinh(x)
    if op == "atanh":                     # elementwise/hyperbolic.py:242-245
        return np.arctanh(x)
    if op == "xexp":                      # elementwise/xexp.py:35-36
        return x * np.exp(x)
    if op == "rel_entr":                  # elementwise/rel_entr.py:36-40
        return rel_entr_scipy(x, v[1])
    if op == "quad_over_lin":             # quad_over_lin.py:39-45
        return np.square(x).sum() / v[1]
    if op == "quad_form":                 # quad_form.py:41-47
        return np.real(np.dot(np.transpose(x), v[1].dot(x)))
    raise NotImplementedError(op)


# ---------------------------------------------------------------------------
# argument verification (the ValueError branches of atoms/atom.py:501-561)
# ---------------------------------------------------------------------------
def _same_var(a, b):
    return a.is_var() and b.is_var() and a.attrs["id"] == b.attrs["id"]


def _verify_jac(node):
    op = node.op
    if op in _ir.ELEMENTWISE_UNARY or op == "quad_form":
        return node.args[0].is_var()       # e.g. elementwise/exp.py:109-110, power.py:424-431
    if op in ("rel_entr",):
        return _verify_hess(node)          # elementwise/rel_entr.py:126-127
    if op == "quad_over_lin":              # quad_over_lin.py:175-176
        return node.args[0].is_var() and node.args[1].is_var()
    if op == "multiply":                   # affine/binary_operators.py:548-549
        return _verify_hess(node)
    if op == "matmul":                     # affine/binary_operators.py:284-297
        xs = {v.attrs["id"] for v in node.args[0].variables()}
        return not any(v.attrs["id"] in xs for v in node.args[1].variables())
    if op == "sum":                        # affine/sum.py:159-163
        return node.attrs["axis"] in (None, 0, 1)
    if op == "reshape":                    # affine/reshape.py:157-158
        return node.attrs["order"] == "F"
    if op == "promote":                    # affine/promote.py:123-124
        return node.args[0].size == 1
    if op == "broadcast_to":               # affine/broadcast_to.py:84-109
        return _broadcast_type(node) is not None or len(node.shape) == 2
    return True                            # affine/affine_atom.py:174-175


def _verify_hess(node):
    op = node.op
    if op in _ir.ELEMENTWISE_UNARY or op == "quad_form":
        return node.args[0].is_var()
    if op == "rel_entr":                   # elementwise/rel_entr.py:105-124
        x, y = node.args
        if not (x.size == 1 or y.size == 1 or x.size == y.size):
            return False
        if not (x.is_var() and y.is_var()):
            return False
        return not _same_var(x, y)
    if op == "quad_over_lin":              # quad_over_lin.py:159-160
        return node.args[0].is_var() and node.args[1].is_var()
    if op == "multiply":                   # affine/binary_operators.py:485-509
        x, y = node.args
        if x.size != y.size:
            return False
        if x.is_constant() and y.is_constant():
            return False
        both = x.is_var() and y.is_var()
        one_const = x.is_constant() or y.is_constant()
        xp = x.op == "promote" and y.is_var()
        yp = y.op == "promote" and x.is_var()
        if not (both or one_const or xp or yp):
            return False
        return not (both and _same_var(x, y))
    if op == "matmul":                     # affine/binary_operators.py:240-259
        X, Y = node.args
        if not X.is_var() and not X.is_constant() and not Y.is_constant():
            return False
        if not Y.is_var() and not Y.is_constant() and not X.is_constant():
            return False
        return not _same_var(X, Y)
    if op == "reshape":                    # affine/reshape.py:163-164
        return node.attrs["order"] == "F"
    if op == "broadcast_to":               # affine/broadcast_to.py:178-179
        return _verify_jac(node)
    return True


def _broadcast_type(node):
    """affine/broadcast_to.py:84-109 (row / col / scalar classification)."""
    if len(node.shape) != 2:
        return None
    m, n = node.shape
    xs = tuple(node.args[0].shape)
    xs = (1,) * (2 - len(xs)) + xs
    kind = None
    if xs[0] == 1 and xs[1] == n:
        kind = "row"
    elif xs[0] == m and xs[1] == 1:
        kind = "col"
    if all(s == 1 for s in xs):
        kind = "scalar"
    return kind


def _dims(node):
    """MulExpression.get_dimensions (affine/binary_operators.py:299-307)."""
    if len(node.shape) == 0:
        return (1, 1)
    if len(node.shape) == 1:
        return (node.shape[0], 1)
    return node.shape


def _val(node, env):
    return numeric(node, env)


# ---------------------------------------------------------------------------
# Jacobians: Atom.jacobian + per-atom _jacobian
# ---------------------------------------------------------------------------
def jacobian(node, env):
    """dict {var_id: (rows, cols, vals)} with block-local F-order indices."""
    if node.op == "var":                   # expressions/variable.py:76-79
        r = np.arange(node.size)
        return {node.attrs["id"]: (r, r, np.ones(node.size))}
    if node.is_constant():                 # atoms/atom.py:504-505, constants/constant.py:273-274
        return {}
    if not _verify_jac(node):              # atoms/atom.py:509-510
        raise ValueError("Argument error in jacobian for atom %s." % node.op)
    return _JAC[node.op](node, env)


def _jac_add(node, env):                   # affine/add_expr.py:190-222
    out, need_sum = {}, []
    for arg in node.args:
        if arg.is_constant():
            continue
        for k, v in jacobian(arg, env).items():
            if k in out:
                for i in range(3):
                    out[k][i].extend(v[i])
                need_sum.append(k)
            else:
                out[k] = tuple(list(np.atleast_1d(v[i])) for i in range(3))
    sizes = {v.attrs["id"]: v.size for v in node.variables()}
    for k in set(need_sum):
        r, c, d = out[k]
        coo = sp.coo_matrix((d, (r, c)), shape=(node.size, sizes[k]))
        coo.sum_duplicates()
        out[k] = (coo.row, coo.col, coo.data)
    return {k: (np.array(r), np.array(c), np.array(d)) for k, (r, c, d) in out.items()}


def _jac_neg(node, env):                   # affine/unary_operators.py:129-136
    return {k: (r, c, -d) for k, (r, c, d) in jacobian(node.args[0], env).items()}


def _jac_sum(node, env):                   # affine/sum.py:165-182
    arg = node.args[0]
    sizes = {v.attrs["id"]: v.size for v in node.variables()}
    out = {}
    for k, (r, c, d) in jacobian(arg, env).items():
        if node.attrs["axis"] is None:
            r = np.zeros(len(c), dtype=int)
        else:
            m, _ = arg.shape
            r = r // m if node.attrs["axis"] == 0 else r % m
        coo = sp.coo_matrix((d, (r, c)), shape=(node.size, sizes[k]))
        coo.sum_duplicates()
        out[k] = (coo.row, coo.col, coo.data)
    return out


def _jac_index(node, env):                 # affine/index.py:127-150
    arg = node.args[0]
    rng = [np.arange(s, (e if e is not None else -1), st) for s, e, st in node.attrs["key"]]
    if len(rng) == 1:
        idx = rng[0]
    elif len(rng) == 2:
        idx = np.add.outer(rng[0], rng[1] * arg.shape[0]).flatten(order="F")
    else:
        raise UnboundLocalError("idx")     # the reference leaves idx undefined for ndim>2
    pos = {val: i for i, val in enumerate(idx)}
    out = {}
    for k, (r, c, d) in jacobian(arg, env).items():
        keep = np.where(np.isin(r, idx))[0]
        rr = np.array([pos[t] for t in r[keep]])
        out[k] = (rr, c[keep], d[keep])
    return out


def _jac_special_index(node, env):         # affine/index.py:264-280
    arg = node.args[0]
    sizes = {v.attrs["id"]: v.size for v in node.variables()}
    sel = np.reshape(node.attrs["select"], node.attrs["select"].size, order="F")
    op = sp.eye_array(arg.size, format="csc")[sel]
    out = {}
    for k, (r, c, d) in jacobian(arg, env).items():
        J = sp.coo_array((d, (r, c)), shape=(arg.size, sizes[k]))
        res = (op @ J).tocoo()
        out[k] = (res.coords[0], res.coords[1], res.data)
    return out


def _jac_reshape(node, env):               # affine/reshape.py:160-161
    return jacobian(node.args[0], env)


def _jac_transpose(node, env):             # affine/transpose.py:126-133
    out = {}
    for k, (r, c, d) in jacobian(node.args[0], env).items():
        mapping = np.arange(node.size).reshape(node.shape, order="F").T.reshape(-1, order="F")
        out[k] = (mapping[r], c, d)
    return out


def _jac_promote(node, env):               # affine/promote.py:126-135
    size = node.size
    out = {}
    for k, (_, c, d) in jacobian(node.args[0], env).items():
        out[k] = (np.repeat(np.arange(size), len(c)), np.tile(c, size), np.tile(d, size))
    return out


def _jac_broadcast(node, env):             # affine/broadcast_to.py:111-176
    m, n = node.shape
    kind = _broadcast_type(node)
    out = {}
    for k, (r, c, d) in jacobian(node.args[0], env).items():
        if kind == "row":
            rr = np.repeat(r * m, m) + np.tile(np.arange(m), len(r))
            out[k] = (rr, np.repeat(c, m), np.repeat(d, m))
        elif kind == "col":
            rr = np.repeat(r, n) + np.tile(np.arange(n) * m, len(r))
            out[k] = (rr, np.repeat(c, n), np.repeat(d, n))
        elif kind == "scalar":
            rr = np.tile(np.arange(m * n), len(r))
            out[k] = (rr, np.repeat(c, m * n), np.repeat(d, m * n))
        else:
            raise NotImplementedError("Jacobian not implemented for broadcast_to.")
    return out


def _jac_multiply(node, env):              # affine/binary_operators.py:552-591
    x, y = node.args
    if x.is_constant():
        xv = _flatF(np.atleast_1d(_dense(_val(x, env))))
        return {k: (r, c, xv[r] * d) for k, (r, c, d) in jacobian(y, env).items()}
    if y.is_constant():
        yv = _flatF(np.atleast_1d(_dense(_val(y, env))))
        return {k: (r, c, yv[r] * d) for k, (r, c, d) in jacobian(x, env).items()}
    if not x.is_var() and x.is_affine():
        xvar = x.args[0]
        idxs = np.arange(y.size, dtype=int)
        return {xvar.attrs["id"]: (idxs, np.zeros(y.size, dtype=int), _val(y, env)),
                y.attrs["id"]: (idxs, idxs, _val(x, env))}
    if not y.is_var() and y.is_affine():
        yvar = y.args[0]
        idxs = np.arange(x.size, dtype=int)
        return {x.attrs["id"]: (idxs, idxs, _val(y, env)),
                yvar.attrs["id"]: (idxs, np.zeros(x.size, dtype=int), _val(x, env))}
    idxs = np.arange(x.size, dtype=int)
    return {x.attrs["id"]: (idxs, idxs, _flatF(_val(y, env))),
            y.attrs["id"]: (idxs, idxs, _flatF(_val(x, env)))}


def _dense(v):
    return v.toarray() if sp.issparse(v) else v


def _jac_matmul(node, env):                # affine/binary_operators.py:309-369
    X, Y = node.args
    m, _ = _dims(X)
    _, p = _dims(Y)
    sizes = {v.attrs["id"]: v.size for v in node.variables()}
    dx_dict, dy_dict = {}, {}
    if not X.is_constant():
        yv = _val(Y, env)
        dx = sp.kron(yv.T, sp.eye(m), format="csr")
        if not X.is_var():
            for k, (r, c, d) in jacobian(X, env).items():
                Jx = sp.coo_array((d, (r, c)), shape=(dx.shape[1], sizes[k])).tocsc()
                Jx = (dx @ Jx).tocoo()
                dx_dict[k] = (Jx.row, Jx.col, Jx.data)
        else:
            dx = dx.tocoo()
            dx_dict = {X.attrs["id"]: (dx.row, dx.col, dx.data)}
    if not Y.is_constant():
        xv = _val(X, env)
        dy = sp.kron(sp.eye(p), xv, format="csr")
        if not Y.is_var():
            for k, (r, c, d) in jacobian(Y, env).items():
                Jy = sp.coo_array((d, (r, c)), shape=(dy.shape[1], sizes[k])).tocsc()
                Jy = (dy @ Jy).tocoo()
                dy_dict[k] = (Jy.row, Jy.col, Jy.data)
        else:
            dy = dy.tocoo()
            dy_dict = {Y.attrs["id"]: (dy.row, dy.col, dy.data)}
    if X.is_constant() and not Y.is_constant():
        return dy_dict
    if not X.is_constant() and Y.is_constant():
        return dx_dict
    dx_dict.update(dy_dict)
    return dx_dict


def _diag_rule(fn):
    def rule(node, env):
        x = node.args[0]
        idxs = np.arange(x.size, dtype=int)
        return {x.attrs["id"]: (idxs, idxs, fn(node, env[x.attrs["id"]]))}
    return rule


def _jac_power(node, env):                 # elementwise/power.py:433-449
    p = node.attrs["p_rational"] if node.attrs["p_rational"] is not None else node.attrs["p"]
    if p == 0:
        return {}
    x = node.args[0]
    idxs = np.arange(x.size, dtype=int)
    vals = float(p) * np.power(_flatF(env[x.attrs["id"]]), float(p) - 1)
    return {x.attrs["id"]: (idxs, idxs, vals)}


def _jac_rel_entr(node, env):              # elementwise/rel_entr.py:129-148
    x, y = node.args
    xv, yv = env[x.attrs["id"]], env[y.attrs["id"]]
    dx = _flatF(np.log(xv / yv)) + 1
    dy = -_flatF(xv / yv)
    z = np.array([0], dtype=int)
    if x.size == 1:
        idxs = np.arange(y.size, dtype=int)
        return {x.attrs["id"]: (z, z, np.array([np.sum(dx)])), y.attrs["id"]: (idxs, idxs, dy)}
    if y.size == 1:
        idxs = np.arange(x.size, dtype=int)
        return {x.attrs["id"]: (idxs, idxs, dx), y.attrs["id"]: (z, z, np.array([np.sum(dy)]))}
    idxs = np.arange(x.size, dtype=int)
    return {x.attrs["id"]: (idxs, idxs, dx), y.attrs["id"]: (idxs, idxs, dy)}


def _jac_quad_over_lin(node, env):         # quad_over_lin.py:178-185
    x, y = node.args
    xv, yv = env[x.attrs["id"]], env[y.attrs["id"]]
    idxs = np.arange(x.size, dtype=int)
    dx = 2.0 * _flatF(xv / yv)
    dy = -np.array([np.sum(xv ** 2) / (yv ** 2)])
    return {x.attrs["id"]: (np.zeros(x.size, dtype=int), idxs, dx),
            y.attrs["id"]: (np.array([0]), np.array([0]), dy)}


def _jac_quad_form(node, env):             # quad_form.py:154-160
    x, Q = node.args
    vals = 2 * (Q.attrs["value"] @ env[x.attrs["id"]]).T
    return {x.attrs["id"]: (np.zeros(x.size, dtype=int), np.arange(x.size, dtype=int), vals)}


_JAC = {
    "add": _jac_add, "neg": _jac_neg, "sum": _jac_sum, "index": _jac_index,
    "special_index": _jac_special_index, "reshape": _jac_reshape,
    "transpose": _jac_transpose, "promote": _jac_promote, "broadcast_to": _jac_broadcast,
    "multiply": _jac_multiply, "matmul": _jac_matmul,
    "exp": _diag_rule(lambda n, x: np.exp(_flatF(x))),                    # exp.py:112-121
    "log": _diag_rule(lambda n, x: 1.0 / _flatF(x)),                      # log.py:118-127
    "entr": _diag_rule(lambda n, x: -np.log(_flatF(x)) - 1),              # entr.py:116-120
    "logistic": _diag_rule(lambda n, x: np.exp(_flatF(x)) / (1 + np.exp(_flatF(x)))),  # logistic.py:108-113
    "power": _jac_power,
    "sin": _diag_rule(lambda n, x: np.cos(_flatF(x))),                    # trig.py:99-103
    "cos": _diag_rule(lambda n, x: -np.sin(_flatF(x))),                   # trig.py:179-183
    "tan": _diag_rule(lambda n, x: 1 / _flatF(np.cos(x)) ** 2),           # trig.py:261-265
    "sinh": _diag_rule(lambda n, x: np.cosh(_flatF(x))),                  # hyperbolic.py:94-98
    "tanh": _diag_rule(lambda n, x: 1 / np.cosh(_flatF(x)) ** 2),         # hyperbolic.py:169-173
    "asinh": _diag_rule(lambda n, x: 1.0 / _flatF(np.sqrt(1.0 + x ** 2))),  # hyperbolic.py:228-232
    "atanh": _diag_rule(lambda n, x: 1.0 / _flatF(1.0 - x ** 2)),         # hyperbolic.py:287-291
    "xexp": _diag_rule(lambda n, x: _flatF(np.exp(x)) * (1 + _flatF(x))),  # xexp.py:108-112
    "rel_entr": _jac_rel_entr, "quad_over_lin": _jac_quad_over_lin, "quad_form": _jac_quad_form,
}


# ---------------------------------------------------------------------------
# Hessian-vector rules: Atom.hess_vec + per-atom _hess_vec
# ---------------------------------------------------------------------------
def hess_vec(node, vec, env):
    """dict {(var_id, var_id): (rows, cols, vals)} = sum_i vec[i] * Hessian(node_i)."""
    if node.op in ("var", "const"):        # variable.py:73-74, constant.py:119-125
        return {}
    if np.size(vec) != node.size:          # atoms/atom.py:546-548
        raise ValueError("Dimension mismatch in hess_vec. vec.size != phi(x).size")
    if node.is_affine():                   # atoms/atom.py:551-552
        return {}
    if not _verify_hess(node):             # atoms/atom.py:556-559
        raise ValueError("Argument error in hess_vec for atom %s." % node.op)
    return _hv_inner(node, vec, env)


def _hv_inner(node, vec, env):
    """The atom-specific ``_hess_vec`` (Promote / broadcast_to call it directly on
    their argument, bypassing the checks above: promote.py:120-121, broadcast_to.py:181-192)."""
    if node.op == "var":
        raise AttributeError("'Variable' object has no attribute '_hess_vec'")
    return _HV[node.op](node, vec, env)


def _hv_add(node, vec, env):               # affine/add_expr.py:149-184
    out, need_sum = {}, []
    for arg in node.args:
        if arg.is_affine():
            continue
        for k, v in hess_vec(arg, vec, env).items():
            if k in out:
                for i in range(3):
                    out[k][i].extend(v[i])
                need_sum.append(k)
            else:
                out[k] = tuple(list(np.atleast_1d(v[i])) for i in range(3))
    sizes = {v.attrs["id"]: v.size for v in node.variables()}
    for k in set(need_sum):
        r, c, d = out[k]
        coo = sp.coo_matrix((d, (r, c)), shape=(sizes[k[0]], sizes[k[0]]))
        coo.sum_duplicates()
        out[k] = (coo.row, coo.col, coo.data)
    return {k: (np.array(r), np.array(c), np.array(d)) for k, (r, c, d) in out.items()}


def _hv_neg(node, vec, env):               # affine/unary_operators.py:122-124
    return hess_vec(node.args[0], -vec, env)


def _hv_sum(node, vec, env):               # affine/sum.py:146-156
    arg = node.args[0]
    if node.attrs["axis"] is None:
        return hess_vec(arg, vec * np.ones(arg.size), env)
    m, n = arg.shape
    rep = np.repeat(vec, m) if node.attrs["axis"] == 0 else np.tile(vec, n)
    return hess_vec(arg, rep, env)


def _hv_index(node, vec, env):             # affine/index.py:120-125 (quirk Q6: flat scatter)
    e = np.zeros(node.args[0].size)
    e[_ir.decode_key(node.attrs["orig_key"])] = vec
    return hess_vec(node.args[0], e, env)


def _hv_special_index(node, vec, env):     # affine/index.py:254-258
    sel = np.reshape(node.attrs["select"], node.attrs["select"].size, order="F")
    e = np.zeros(node.args[0].size)
    e[sel] = vec
    return hess_vec(node.args[0], e, env)


def _hv_reshape(node, vec, env):           # affine/reshape.py:166-167
    return hess_vec(node.args[0], vec, env)


def _hv_transpose(node, vec, env):         # affine/transpose.py:119-123
    return hess_vec(node.args[0], vec.reshape(node.shape, order="F").T.reshape(-1, order="F"), env)


def _hv_promote(node, vec, env):           # affine/promote.py:120-121
    return _hv_inner(node.args[0], np.sum(vec), env)


def _hv_broadcast(node, vec, env):         # affine/broadcast_to.py:181-192
    m, n = node.shape
    kind = _broadcast_type(node)
    if env.get("_broadcast_types_unset") and node.args[0].op == "broadcast_to":
        # broadcast_to caches its type lazily in _verify_jacobian_args (broadcast_to.py:84-110), i.e. when the
        # jacobian()/hess_vec() WRAPPER visits it; the rule below calls the child's _hess_vec directly, so an inner
        # broadcast_to the wrapper has never seen still has type None and falls through to NotImplementedError
        # (broadcast_to.py:192).  That is the state of the OBJECTIVE at the structure pass: its Jacobian is only taken
        # in gradient() (nlp_solver.py:218-235), after hessianstructure().
        raise NotImplementedError("hess-vec not implemented for broadcast_to.")
    if kind == "row":
        return _hv_inner(node.args[0], vec.reshape(n, m).sum(axis=1), env)
    if kind == "col":
        return _hv_inner(node.args[0], vec.reshape(n, m).sum(axis=0), env)
    if kind == "scalar":
        return _hv_inner(node.args[0], vec.sum(), env)
    raise NotImplementedError("hess-vec not implemented for broadcast_to.")


def _hv_multiply(node, vec, env):          # affine/binary_operators.py:511-546
    x, y = node.args
    if x.is_constant():
        return hess_vec(y, _flatF(_dense(_val(x, env))) * vec, env)
    if y.is_constant():
        return hess_vec(x, _flatF(_dense(_val(y, env))) * vec, env)
    if not x.is_var() and x.is_affine():
        xvar = x.args[0]
        z = np.zeros(xvar.size, dtype=int)
        c = np.arange(y.size, dtype=int)
        return {(xvar.attrs["id"], y.attrs["id"]): (z, c, vec),
                (y.attrs["id"], xvar.attrs["id"]): (c, z, vec)}
    if not y.is_var() and y.is_affine():
        yvar = y.args[0]
        z = np.zeros(yvar.size, dtype=int)
        c = np.arange(x.size, dtype=int)
        return {(x.attrs["id"], yvar.attrs["id"]): (c, z, vec),
                (yvar.attrs["id"], x.attrs["id"]): (z, c, vec)}
    r = np.arange(x.size, dtype=int)
    return {(x.attrs["id"], y.attrs["id"]): (r, r, vec),
            (y.attrs["id"], x.attrs["id"]): (r, r, vec)}


def _hv_matmul(node, vec, env):            # affine/binary_operators.py:261-282
    X, Y = node.args
    m, n = _dims(X)
    _, p = _dims(Y)
    if X.is_constant():
        B = _val(X, env).T @ np.reshape(vec, (m, p), order="F")
        return hess_vec(Y, _flatF(B), env)
    if Y.is_constant():
        B = np.reshape(vec, (m, p), order="F") @ _val(Y, env).T
        return hess_vec(X, _flatF(B), env)
    rows = np.tile(np.arange(m * n), p)
    cols = np.repeat(np.arange(n * p), m)
    vals = vec[(cols // n) * m + (rows % m)]
    return {(X.attrs["id"], Y.attrs["id"]): (rows, cols, vals),
            (Y.attrs["id"], X.attrs["id"]): (cols, rows, vals)}


def _diag_hv(fn):
    def rule(node, vec, env):
        x = node.args[0]
        idxs = np.arange(x.size, dtype=int)
        return {(x.attrs["id"], x.attrs["id"]): (idxs, idxs, fn(env[x.attrs["id"]], vec))}
    return rule


def _hv_power(node, vec, env):             # elementwise/power.py:408-422
    p = node.attrs["p_rational"] if node.attrs["p_rational"] is not None else node.attrs["p"]
    if p == 0 or p == 1:
        return {}
    x = node.args[0]
    hv = float(p) * float(p - 1) * np.power(_flatF(env[x.attrs["id"]]), float(p) - 2)
    idxs = np.arange(x.size, dtype=int)
    return {(x.attrs["id"], x.attrs["id"]): (idxs, idxs, hv * vec)}


def _logistic_hv(x, vec):                  # elementwise/logistic.py:97-103
    e = np.exp(_flatF(x))
    return e / (e + 1) ** 2 * vec


def _hv_rel_entr(node, vec, env):          # elementwise/rel_entr.py:150-180
    x, y = node.args
    xi, yi = x.attrs["id"], y.attrs["id"]
    xv, yv = env[xi], env[yi]
    dx2 = vec / _flatF(xv)
    dy2 = vec * _flatF(xv / (yv ** 2))
    dxdy = -vec / _flatF(yv)
    z1 = np.array([0], dtype=int)
    if x.size == 1:
        idxs = np.arange(y.size, dtype=int)
        zy = np.zeros(y.size, dtype=int)
        return {(xi, xi): (z1, z1, np.array([np.sum(dx2)])), (yi, yi): (idxs, idxs, dy2),
                (xi, yi): (zy, idxs, dxdy), (yi, xi): (idxs, zy, dxdy)}
    if y.size == 1:
        idxs = np.arange(x.size, dtype=int)
        zx = np.zeros(x.size, dtype=int)
        return {(xi, xi): (idxs, idxs, dx2), (yi, yi): (z1, z1, np.array([np.sum(dy2)])),
                (xi, yi): (idxs, zx, dxdy), (yi, xi): (zx, idxs, dxdy)}
    idxs = np.arange(x.size, dtype=int)
    return {(xi, xi): (idxs, idxs, dx2), (yi, yi): (idxs, idxs, dy2),
            (xi, yi): (idxs, idxs, dxdy), (yi, xi): (idxs, idxs, dxdy)}


def _hv_quad_over_lin(node, vec, env):     # quad_over_lin.py:162-173
    x, y = node.args
    xi, yi = x.attrs["id"], y.attrs["id"]
    xv, yv = env[xi], env[yi]
    idxs = np.arange(x.size, dtype=int)
    zx = np.zeros(x.size, dtype=int)
    dx2 = vec * (2.0 * np.ones(x.size) / yv)
    dy2 = vec * 2.0 * (np.sum(xv ** 2) / (yv ** 3))
    dxdy = vec * -_flatF(2.0 * xv / (yv ** 2))
    return {(xi, xi): (idxs, idxs, dx2), (yi, yi): (np.array([0]), np.array([0]), dy2),
            (xi, yi): (idxs, zx, dxdy), (yi, xi): (zx, idxs, dxdy)}


def _hv_quad_form(node, vec, env):         # quad_form.py:143-149
    x, Q = node.args
    Qc = sp.coo_matrix(Q.attrs["value"])
    return {(x.attrs["id"], x.attrs["id"]): (Qc.row, Qc.col, 2 * vec * Qc.data)}


_HV = {
    "add": _hv_add, "neg": _hv_neg, "sum": _hv_sum, "index": _hv_index,
    "special_index": _hv_special_index, "reshape": _hv_reshape, "transpose": _hv_transpose,
    "promote": _hv_promote, "broadcast_to": _hv_broadcast,
    "multiply": _hv_multiply, "matmul": _hv_matmul,
    "exp": _diag_hv(lambda x, v: np.exp(_flatF(x)) * v),                     # exp.py:102-107
    "log": _diag_hv(lambda x, v: -v / (_flatF(x) ** 2)),                     # log.py:108-113
    "entr": _diag_hv(lambda x, v: -v / _flatF(x)),                           # entr.py:106-111
    "logistic": _diag_hv(_logistic_hv),
    "power": _hv_power,
    "sin": _diag_hv(lambda x, v: -np.sin(_flatF(x)) * v),                    # trig.py:90-94
    "cos": _diag_hv(lambda x, v: -np.cos(_flatF(x)) * v),                    # trig.py:170-174
    "tan": _diag_hv(lambda x, v: _flatF(2 * np.tan(x) / np.cos(x) ** 2) * v),  # trig.py:251-256
    "sinh": _diag_hv(lambda x, v: np.sinh(_flatF(x)) * v),                   # hyperbolic.py:85-89
    "tanh": _diag_hv(lambda x, v: -2 * _flatF(np.tanh(x) / np.cosh(x) ** 2) * v),  # hyperbolic.py:160-164
    "asinh": _diag_hv(lambda x, v: _flatF(-x / (1.0 + x ** 2) ** 1.5) * v),  # hyperbolic.py:219-223
    "atanh": _diag_hv(lambda x, v: _flatF(2.0 * x / (1.0 - x ** 2) ** 2) * v),  # hyperbolic.py:278-282
    "xexp": _diag_hv(lambda x, v: np.exp(_flatF(x)) * (2 + _flatF(x)) * v),  # xexp.py:117-121
    "rel_entr": _hv_rel_entr, "quad_over_lin": _hv_quad_over_lin, "quad_form": _hv_quad_form,
}


# ---------------------------------------------------------------------------
# the seven callbacks  (reductions/solvers/nlp_solvers/nlp_solver.py:181-427)
# ---------------------------------------------------------------------------
def _sampled(res):
    """``M[rows, cols].data`` (nlp_solver.py:275,371): an np.matrix normally, a sparse
    matrix with empty ``data`` when the index lists are empty."""
    if sp.issparse(res):
        return np.asarray(res.data, dtype=np.float64).ravel()
    return np.asarray(res, dtype=np.float64).ravel()


class RefOracles:
    """Array-based restatement of ``Oracles``.

    The reference accumulates triplets in Python lists; here they are NumPy
    arrays concatenated in the same order, and the same SciPy calls
    (``coo.sum_duplicates``, ``csr_matrix`` sampling) finish the job, so results
    are identical while the CPU baseline stays a fair (faster) one.
    """

    def __init__(self, prob):
        self.prob = prob
        self.n, self.m = prob.n, prob.m
        self.vars = prob.variables
        self.grad_obj = np.zeros(self.n, dtype=np.float64)      # nlp_solver.py:184
        self.has_jac_structure = False
        self.has_hess_structure = False
        self.has_affine_cache = False
        self.affine_coo = None
        self.iterations = 0
        self._affine = [c.is_affine() for c in prob.constraints]

    def _env(self, x):                     # set_variable_value, nlp_solver.py:205-210
        env, off = {}, 0
        for v in self.vars:
            env[v.attrs["id"]] = x[off:off + v.size].reshape(v.shape, order="F")
            off += v.size
        return env

    def objective(self, x):                # nlp_solver.py:212-216
        return numeric(self.prob.objective, self._env(x))

    def gradient(self, x):                 # nlp_solver.py:218-235 (scatter-assign, reused buffer)
        env = self._env(x)
        self.grad_obj.fill(0)
        gd = jacobian(self.prob.objective, env)
        self._objective_jacobian_taken = True      # the wrappers have now visited every node of the objective
        off = 0
        for v in self.vars:
            if v.attrs["id"] in gd:
                _, cols, vals = gd[v.attrs["id"]]
                self.grad_obj[off + cols] = vals
            off += v.size
        return self.grad_obj

    def constraints(self, x):              # nlp_solver.py:237-244
        env = self._env(x)
        return np.concatenate([_flatF(_dense(numeric(c, env))) for c in self.prob.constraints])

    def _jac_triplets(self, env):          # nlp_solver.py:246-265, 278-299
        if self.has_affine_cache:
            R, C, V = [self.affine_coo[0]], [self.affine_coo[1]], [self.affine_coo[2]]
        else:
            R, C, V = [], [], []
        aR, aC, aV = [], [], []
        coff = 0
        for con, aff in zip(self.prob.constraints, self._affine):
            if aff and self.has_affine_cache:
                coff += con.size
                continue
            gd = jacobian(con, env)
            voff = 0
            for v in self.vars:
                if v.attrs["id"] in gd:
                    r, c, d = gd[v.attrs["id"]]
                    r, c = np.asarray(r) + coff, np.asarray(c) + voff
                    d = np.asarray(d, dtype=np.float64).reshape(-1)
                    R.append(r), C.append(c), V.append(d)
                    if aff:
                        aR.append(r), aC.append(c), aV.append(d)
                voff += v.size
            coff += con.size
        cat = lambda L, dt: np.concatenate(L).astype(dt) if L else np.zeros(0, dt)  # noqa: E731
        return (cat(R, np.int64), cat(C, np.int64), cat(V, np.float64)), \
               (cat(aR, np.int64), cat(aC, np.int64), cat(aV, np.float64))

    def jacobian(self, x):                 # nlp_solver.py:278-307
        (r, c, v), _ = self._jac_triplets(self._env(x))
        if not self.has_jac_structure:
            return v
        if not self.permutation_needed:    # nlp_solver.py:270-271
            return v
        J = sp.csr_matrix((v, (r, c)), shape=(self.m, self.n))     # nlp_solver.py:274-276
        return _sampled(J[self.jac_rows, self.jac_cols])

    def jacobianstructure(self):           # nlp_solver.py:309-335
        if self.has_jac_structure:
            return self.jac_rows, self.jac_cols
        x = np.nan * np.ones(self.n)
        (r, c, _), aff = self._jac_triplets(self._env(x))
        self.affine_coo = aff
        self.has_jac_structure = True
        self.has_affine_cache = True
        self.permutation_needed = not all(self._affine)
        self.jac_rows, self.jac_cols = r.astype(np.int32), c.astype(np.int32)
        return self.jac_rows, self.jac_cols

    def _hess_coo(self, x, duals, obj_factor):   # nlp_solver.py:337-364, 394-413
        env = self._env(x)
        R, C, V = [], [], []
        offs, off = {}, 0
        for v in self.vars:
            offs[v.attrs["id"]] = off
            off += v.size

        def parse(hd):                     # parse_hess_dict: var1-major, var2-minor block order
            for v1 in self.vars:
                for v2 in self.vars:
                    key = (v1.attrs["id"], v2.attrs["id"])
                    if key in hd:
                        r, c, d = hd[key]
                        R.append(np.asarray(r) + offs[key[0]])
                        C.append(np.asarray(c) + offs[key[1]])
                        V.append(np.asarray(d, dtype=np.float64).reshape(-1))

        env["_broadcast_types_unset"] = not getattr(self, "_objective_jacobian_taken", False)
        try:
            parse(hess_vec(self.prob.objective, np.array([obj_factor]), env))
        finally:
            env["_broadcast_types_unset"] = False
        coff = 0
        for con in self.prob.constraints:
            parse(hess_vec(con, duals[coff:coff + con.size], env))
            coff += con.size
        cat = lambda L, dt: np.concatenate(L).astype(dt) if L else np.zeros(0, dt)  # noqa: E731
        coo = sp.coo_matrix((cat(V, np.float64), (cat(R, np.int64), cat(C, np.int64))),
                            shape=(self.n, self.n))
        coo.sum_duplicates()               # sum_coo, nlp_solver.py:359-364
        return coo.row, coo.col, coo.data

    def hessian(self, x, duals, obj_factor):     # nlp_solver.py:394-421
        r, c, v = self._hess_coo(x, np.asarray(duals, dtype=np.float64), obj_factor)
        if not self.has_hess_structure:
            return v
        H = sp.csr_matrix((v, (r, c)), shape=(self.n, self.n))     # nlp_solver.py:366-372
        return _sampled(H[self.hess_rows, self.hess_cols])

    def hessianstructure(self):            # nlp_solver.py:374-392
        if self.has_hess_structure:
            return self.hess_rows, self.hess_cols
        x = np.nan * np.ones(self.n)
        r, c, _ = self._hess_coo(x, np.ones(self.m), 1.0)
        self.has_hess_structure = True
        mask = r >= c
        self.hess_rows, self.hess_cols = r[mask].astype(np.int32), c[mask].astype(np.int32)
        return self.hess_rows, self.hess_cols

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.iterations = iter_count       # nlp_solver.py:423-427
