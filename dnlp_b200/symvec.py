"""Symbolic sparse-polynomial vectors: the value algebra of the DAG compiler.

A ``SymVec`` is a length-``K`` vector whose entry ``k`` is

    sum over terms t with row[t] == k of   coef[t] * V[f1[t]] * V[f2[t]]

where ``V`` is the device value buffer (slots: the point x, the multipliers
(sigma, lambda), and outputs of earlier tape instructions) and a factor index of
``-1`` means "1".  Constants are terms with no factors.  Every derivative rule of
the reference manipulates triplet *values* only through a handful of operations
(negate, scale by a constant, gather/permute, sum duplicates, multiply by another
value vector); those are exactly the methods below, all NumPy-vectorised so that
50 M-term vectors compile in seconds.

Products that would exceed two factors per term, and nonlinear functions of a
non-trivial entry, are handled by *materialising* a SymVec: the compiler emits a
tape instruction that evaluates it into fresh slots and continues with a SymVec of
bare slots (``Builder.materialise``).
"""
import numpy as np
import scipy.sparse as sp

NONE = -1


def _i64(a):
    return np.asarray(a, dtype=np.int64)


def _i32(a):
    """Slot / entry indices are stored as int32 (value buffers stay below 2^31 slots) to halve the
    compiler's memory footprint on 50 M-term vectors; index *arithmetic* is done in int64."""
    return np.asarray(a, dtype=np.int32)


def stable_order(key, K):
    """``np.argsort(key, kind="stable")`` for integer keys in [0, K).  Large inputs go through SciPy's
    COO -> CSR conversion, a C counting sort that keeps the input order inside a row (3-4x faster
    than NumPy's merge sort on the 50 M-term vectors of config 5)."""
    n = key.size
    if n < (1 << 16) or K > 4 * n or n >= 2 ** 31 - 1 or K >= 2 ** 31 - 1:
        return np.argsort(key, kind="stable")
    m = sp.coo_matrix((np.ones(n, dtype=np.int8), (key, np.arange(n, dtype=np.int32))), shape=(int(K), n))
    return m.tocsr().indices.astype(np.int64)      # column = original position, ascending inside a row


def _sorted_by_row(K, row, coef, f1, f2):
    """SymVec with terms stably ordered by entry; skips the sort when already ordered."""
    if row.size > 1 and np.any(row[1:] < row[:-1]):
        order = stable_order(row, K)
        return SymVec(K, row[order], coef[order], f1[order], f2[order])
    return SymVec(K, row, coef, f1, f2)


class SymVec:
    __slots__ = ("K", "row", "coef", "f1", "f2", "_ptr")

    def __init__(self, K, row, coef, f1, f2):
        self.K = int(K)
        self.row = _i32(row)
        self.coef = np.asarray(coef, dtype=np.float64)
        self.f1 = _i32(f1)
        self.f2 = _i32(f2)
        self._ptr = None

    # ---- constructors -----------------------------------------------------
    @staticmethod
    def const(values):
        v = np.asarray(values, dtype=np.float64).reshape(-1)
        k = v.size
        none = np.full(k, NONE, dtype=np.int64)
        return SymVec(k, np.arange(k), v, none, none)

    @staticmethod
    def zeros(K):
        e = np.zeros(0, dtype=np.int64)
        return SymVec(K, e, np.zeros(0), e, e)

    @staticmethod
    def slots(idx):
        idx = _i64(idx).reshape(-1)
        k = idx.size
        return SymVec(k, np.arange(k), np.ones(k), idx, np.full(k, NONE, dtype=np.int64))

    @staticmethod
    def slot_range(start, count):
        return SymVec.slots(np.arange(start, start + count, dtype=np.int64))

    # ---- structure --------------------------------------------------------
    @property
    def nterms(self):
        return self.row.size

    @property
    def ptr(self):
        if self._ptr is None:
            cnt = np.bincount(self.row, minlength=self.K) if self.row.size else np.zeros(self.K, np.int64)
            p = np.zeros(self.K + 1, dtype=np.int64)
            np.cumsum(cnt, out=p[1:])
            self._ptr = p
        return self._ptr

    def term_counts(self):
        return np.diff(self.ptr)

    def is_unit(self):
        """Exactly one term per entry, stored in entry order (slots, constants, simple products)."""
        if self.nterms != self.K:
            return False
        if self.K == 0:
            return True
        return bool(self.row[0] == 0 and self.row[-1] == self.K - 1 and np.all(np.diff(self.row) == 1))

    def is_const_mask(self):
        """Per entry: True when no term depends on a slot (value known at compile time)."""
        dep = np.zeros(self.K, dtype=bool)
        if self.row.size:
            dep[self.row[(self.f1 != NONE) | (self.f2 != NONE)]] = True
        return ~dep

    def const_values(self):
        """Numeric value of the constant part of every entry (terms without factors)."""
        m = (self.f1 == NONE) & (self.f2 == NONE)
        if not m.any():
            return np.zeros(self.K)
        return np.bincount(self.row[m], weights=self.coef[m], minlength=self.K)

    def bare_slots(self):
        """If every entry is exactly ``1.0 * V[s]`` return the slot array, else None."""
        if self.nterms != self.K or self.K == 0:
            return None if self.K else np.zeros(0, np.int64)
        if not np.array_equal(self.row, np.arange(self.K)):
            return None
        if np.any(self.f2 != NONE) or np.any(self.f1 == NONE) or np.any(self.coef != 1.0):
            return None
        return self.f1

    def contiguous_start(self):
        s = self.bare_slots()
        if s is None or s.size == 0:
            return None
        if s.size == 1 or np.array_equal(s, np.arange(s[0], s[0] + s.size)):
            return int(s[0])
        return None

    # ---- linear operations --------------------------------------------------
    def neg(self):
        return SymVec(self.K, self.row, -self.coef, self.f1, self.f2)

    def scale(self, c):
        """Entry k multiplied by the constant c[k] (c scalar or length K)."""
        c = np.asarray(c, dtype=np.float64)
        if c.ndim == 0 or c.size == 1:
            return SymVec(self.K, self.row, self.coef * c.reshape(-1)[0], self.f1, self.f2)
        c = c.reshape(-1)
        assert c.size == self.K, (c.size, self.K)
        return SymVec(self.K, self.row, self.coef * c[self.row], self.f1, self.f2)

    def gather(self, idx):
        """New entry j = old entry idx[j] (selection, permutation or duplication)."""
        idx = _i64(idx).reshape(-1)
        if self.is_unit():
            return SymVec(idx.size, np.arange(idx.size, dtype=np.int32), self.coef[idx], self.f1[idx], self.f2[idx])
        p = self.ptr
        cnt = p[idx + 1] - p[idx]
        total = int(cnt.sum())
        new_row = np.repeat(np.arange(idx.size, dtype=np.int64), cnt)
        if total == 0:
            return SymVec.zeros(idx.size)
        start = np.repeat(p[idx], cnt)
        out_ptr = np.zeros(idx.size + 1, dtype=np.int64)
        np.cumsum(cnt, out=out_ptr[1:])
        within = np.arange(total, dtype=np.int64) - np.repeat(out_ptr[:-1], cnt)
        src = start + within
        return SymVec(idx.size, new_row, self.coef[src], self.f1[src], self.f2[src])

    def scatter_into(self, K, pos):
        """Vector of length K with entry pos[j] = self[j] (pos unique); other entries 0."""
        pos = _i64(pos).reshape(-1)
        assert pos.size == self.K
        new_row = pos[self.row]
        return _sorted_by_row(K, new_row, self.coef, self.f1, self.f2)

    def group_sum(self, group, G):
        """New entry g = sum of old entries k with group[k] == g."""
        group = _i64(group).reshape(-1)
        assert group.size == self.K
        new_row = group[self.row]
        return _sorted_by_row(G, new_row, self.coef, self.f1, self.f2)

    def sum_all(self):
        return self.group_sum(np.zeros(self.K, dtype=np.int64), 1)

    @staticmethod
    def concat(parts):
        parts = list(parts)
        if not parts:
            return SymVec.zeros(0)
        if len(parts) == 1:
            return parts[0]
        offs = [0]
        for p in parts:
            offs.append(offs[-1] + p.K)
        if offs[-1] >= 2 ** 31:
            raise ValueError("vector of %d entries exceeds the int32 index range" % offs[-1])
        row = np.empty(sum(p.nterms for p in parts), dtype=np.int32)
        at = 0
        for p, o in zip(parts, offs[:-1]):               # one pass per part, no int64 round trip
            np.add(p.row, np.int32(o), out=row[at:at + p.nterms])
            at += p.nterms
        return SymVec(offs[-1], row,
                      np.concatenate([p.coef for p in parts]),
                      np.concatenate([p.f1 for p in parts]),
                      np.concatenate([p.f2 for p in parts]))

    def add(self, other):
        assert self.K == other.K
        return _sorted_by_row(self.K, np.concatenate([self.row, other.row]),
                              np.concatenate([self.coef, other.coef]),
                              np.concatenate([self.f1, other.f1]),
                              np.concatenate([self.f2, other.f2]))

    @staticmethod
    def add_many(parts):
        """``parts[0].add(parts[1]).add(parts[2])...`` in one pass: a stable sort of the whole concatenation leaves
        the terms of an entry in part order, each part's own order kept - exactly what the chain of pairwise
        stable sorts produces, without re-sorting the growing prefix once per part."""
        parts = list(parts)
        if len(parts) == 1:
            return parts[0]
        K = parts[0].K
        assert all(p.K == K for p in parts)
        return _sorted_by_row(K, np.concatenate([p.row for p in parts]), np.concatenate([p.coef for p in parts]),
                              np.concatenate([p.f1 for p in parts]), np.concatenate([p.f2 for p in parts]))

    def linear_map(self, out_rows, in_idx, weights, n_out):
        """sum_j M[i, j] * self[j] for a constant sparse M given as COO (out_rows, in_idx, weights)."""
        g = self.gather(in_idx).scale(weights)
        return g.group_sum(out_rows, n_out)

    # ---- products -----------------------------------------------------------
    def arity(self):
        return (self.f1 != NONE).astype(np.int64) + (self.f2 != NONE).astype(np.int64)

    def can_multiply_directly(self, other):
        """Entrywise product stays within two factors per term without expansion blow-up."""
        if not (np.all(self.term_counts() <= 1) and np.all(other.term_counts() <= 1)):
            return False
        a = np.zeros(self.K, dtype=np.int64)
        a[self.row] = self.arity()
        b = np.zeros(other.K, dtype=np.int64)
        b[other.row] = other.arity()
        return bool(np.all(a + b <= 2))

    def mul_simple(self, other):
        """Entrywise product; both operands have at most one term per entry and the
        combined factor count is at most two (see ``can_multiply_directly``)."""
        assert self.K == other.K
        ta = np.full(self.K, -1, dtype=np.int64)
        ta[self.row] = np.arange(self.nterms)
        tb = np.full(other.K, -1, dtype=np.int64)
        tb[other.row] = np.arange(other.nterms)
        rows = np.where((ta >= 0) & (tb >= 0))[0]
        ia, ib = ta[rows], tb[rows]
        coef = self.coef[ia] * other.coef[ib]
        fa1, fa2 = self.f1[ia], self.f2[ia]
        fb1, fb2 = other.f1[ib], other.f2[ib]
        # collect the (at most two) present factors: sort descending so NONE (-1) goes last
        F = -np.sort(-np.stack([fa1, fa2, fb1, fb2], axis=1), axis=1)
        assert not F.size or np.all(F[:, 2] == NONE), "product exceeds two factors per term"
        return SymVec(self.K, rows, coef, F[:, 0], F[:, 1])

    def distribute(self, other):
        """Entrywise product of ``self`` (at most ONE term per entry, at most one factor) with ``other``
        (any number of single-factor terms per entry): (c V[s]) * sum_t c_t V[f_t] = sum_t c c_t V[s] V[f_t].
        The adjoint weights A' lambda of a second-derivative rule thus never have to be written out: the
        Hessian entry phi''_j * sum_i A_ij lambda_i becomes one row of two-factor terms."""
        assert self.K == other.K
        ta = np.full(self.K, -1, dtype=np.int64)
        ta[self.row] = np.arange(self.nterms)
        ia = ta[other.row]
        keep = ia >= 0
        ia = ia[keep]
        coef = self.coef[ia] * other.coef[keep]
        F = -np.sort(-np.stack([self.f1[ia], self.f2[ia], other.f1[keep], other.f2[keep]], axis=1), axis=1)
        assert not F.size or np.all(F[:, 2] == NONE), "product exceeds two factors per term"
        return SymVec(self.K, other.row[keep], coef, F[:, 0], F[:, 1])

    # ---- clean-up -----------------------------------------------------------
    def drop_zero_constants(self):
        """Remove constant terms that are exactly zero (e.g. the `- 0` right-hand side that
        lower_equality appends to every row); linear time, safe on 50 M-term vectors."""
        z = (self.f1 == NONE) & (self.f2 == NONE) & (self.coef == 0.0)
        if not z.any():
            return self
        k = ~z
        return SymVec(self.K, self.row[k], self.coef[k], self.f1[k], self.f2[k])

    def simplify(self):
        """Merge like terms (same entry, same factor pair) and drop exact-zero coefficients
        of factor terms.  Constant terms are kept even when zero-valued only if the entry
        has no other term (so an explicit structural zero stays representable)."""
        if self.nterms == 0:
            return self
        lo = np.minimum(self.f1, self.f2)
        hi = np.maximum(self.f1, self.f2)
        order = np.lexsort((lo, hi, self.row))
        r, l, h, c = self.row[order], lo[order], hi[order], self.coef[order]
        new = np.ones(r.size, dtype=bool)
        new[1:] = (r[1:] != r[:-1]) | (l[1:] != l[:-1]) | (h[1:] != h[:-1])
        starts = np.nonzero(new)[0]
        coef = np.add.reduceat(c, starts)
        r, l, h = r[starts], l[starts], h[starts]
        keep = (coef != 0.0) | np.isnan(coef)
        # factors: put the present one first
        return SymVec(self.K, r[keep], coef[keep], h[keep], l[keep])
