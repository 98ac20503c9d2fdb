"""Compile-once reuse of the GPU oracle across solves of the same smooth problem.

The reference's ``best_of=N`` loop (cvxpy/problems/problem.py:1249-1275, 1643-1693) re-samples the
variables and re-applies the whole reduction chain for every start, so ``NLPsolver.apply`` builds a
fresh ``Oracles`` object each time (nlp_solver.py:61-79).  For the GPU oracle that would mean
re-running the DAG compiler and re-uploading the tape (10 s and 1 GB for the dense n = 8192
eigen-QCQP) although only the initial point differs.  ``fingerprint`` identifies a smooth problem
by everything the tape depends on - the expression DAG, every constant's bytes, the variable
layout - and by nothing else (variable ids, names, values and bounds do not enter the tape), and
``OracleCache`` hands back the already-compiled oracle when the fingerprint matches.
"""
import hashlib
from collections import OrderedDict
from fractions import Fraction

import numpy as np
import scipy.sparse as sp

_VAR_ATTRS_IGNORED = ("id", "name", "value", "lb", "ub")


def _feed_array(h, a):
    a = np.ascontiguousarray(a)
    h.update(str((a.dtype.str, a.shape)).encode())
    if a.size:                                   # (memoryview.cast rejects shapes with a zero in them)
        h.update(a.reshape(-1).view(np.uint8))


def _feed_value(h, v):
    if sp.issparse(v):
        # the stored format and the raw index order enter the digest: the rules emit triplets in the
        # constant's own COO order (rules.py `_tag_matrix`), so two constants with the same entries in a
        # different storage order produce differently ordered patterns
        h.update(("sparse:%s%s" % (v.format, v.shape)).encode())
        if v.format in ("csr", "csc", "bsr"):
            _feed_array(h, v.indptr)
            _feed_array(h, v.indices)
        elif v.format == "coo":
            _feed_array(h, v.coords[0] if hasattr(v, "coords") else v.row)
            _feed_array(h, v.coords[1] if hasattr(v, "coords") else v.col)
        else:
            c = sp.coo_array(v)
            _feed_array(h, c.coords[0])
            _feed_array(h, c.coords[1])
            v = c
        _feed_array(h, np.asarray(v.data, np.float64))
    elif isinstance(v, np.ndarray):
        _feed_array(h, v)
    elif isinstance(v, Fraction):
        h.update(("frac%d/%d" % (v.numerator, v.denominator)).encode())
    elif isinstance(v, (list, tuple)):
        h.update(b"seq%d" % len(v))
        for e in v:
            _feed_value(h, e)
    elif isinstance(v, dict):
        h.update(b"map%d" % len(v))
        for k in sorted(v):
            h.update(str(k).encode())
            _feed_value(h, v[k])
    elif isinstance(v, float):
        h.update(np.float64(v).tobytes())
    else:
        h.update(repr(v).encode())


def fingerprint(prob):
    """Hex digest identifying the tape ``compile_problem(prob)`` would produce."""
    h = hashlib.blake2b(digest_size=20)
    var_index = {id(v): i for i, v in enumerate(prob.variables)}
    param_index = {id(q): i for i, q in enumerate(getattr(prob, "params", []))}
    memo = {}

    def visit(n):
        k = id(n)
        if k in memo:
            return memo[k]
        kids = [visit(a) for a in n.args]
        idx = len(memo)
        h.update(("|%d:%s%s<%s>" % (idx, n.op, n.shape, ",".join(map(str, kids)))).encode())
        if n.op == "var":
            # a variable is its position in the flat layout; one that is not listed cannot be compiled
            h.update(b"v%d" % var_index.get(k, -1))
        elif n.op == "param":
            # a parameter is a slot: its position and shape enter the tape, its VALUE does not
            h.update(b"p%d" % param_index.get(k, -1))
        else:
            for name in sorted(n.attrs):
                h.update(name.encode())
                _feed_value(h, n.attrs[name])
        memo[k] = idx
        return idx

    roots = [visit(prob.objective)] + [visit(c) for c in prob.constraints] + [visit(v) for v in prob.variables]
    h.update(("roots%s n%d m%d" % (roots, prob.n, prob.m)).encode())
    return h.hexdigest()


class OracleCache:
    """Small LRU of live oracles keyed by ``fingerprint`` (plus whatever the caller adds: device,
    build options).  Evicted oracles are only DROPPED, never closed here: the caller may still hold
    them (``data['oracles']`` and the bound callbacks of ``get_problem_data``), and an oracle frees its
    HBM by itself when the last reference goes away (``GpuOracles.__del__``)."""

    def __init__(self, capacity=2):
        self.capacity = int(capacity)
        self._items = OrderedDict()
        self.hits = self.misses = 0

    def get(self, prob, build, extra_key=()):
        """``build(prob)`` -> oracle is called on a miss.  Returns (oracle, hit)."""
        if self.capacity <= 0:
            self.misses += 1
            return build(prob), False
        key = (fingerprint(prob),) + tuple(extra_key)
        if key in self._items:
            self._items.move_to_end(key)
            self.hits += 1
            return self._items[key], True
        self.misses += 1
        o = build(prob)
        self._items[key] = o
        while len(self._items) > self.capacity:
            self._items.popitem(last=False)
        return o, False

    def clear(self):
        self._items.clear()
