"""Unions of half-open integer intervals (lo, hi] with set algebra by endpoint sweeps.

Replaces the `portion` library the reference uses on its 1 ms frame grid (analysis/preprocess.py:33, analysis/analyse.py:118,
analysis/utils.py:27-37): `P.openclosed(a, b)` with integer bounds, `|`, `&`, `-`, `contains`, `overlaps`, and the length
`len(list(P.iterate(x, step=1)))` = the number of integer points = sum(hi - lo).  Intervals are kept sorted, disjoint and
merged when they touch ((1, 3] | (3, 5] = (1, 5], as in portion), so every operation is one sort of the endpoints."""
import numpy as np


class IntervalSet:
    __slots__ = ("lo", "hi")

    def __init__(self, lo=None, hi=None, _normalised=False):
        lo = np.zeros(0, np.int64) if lo is None else np.asarray(lo, dtype=np.int64).reshape(-1)
        hi = np.zeros(0, np.int64) if hi is None else np.asarray(hi, dtype=np.int64).reshape(-1)
        if lo.shape != hi.shape:
            raise ValueError("lo and hi differ in length")
        if not _normalised:
            keep = hi > lo                       # (a, a] and reversed bounds are empty
            lo, hi = lo[keep], hi[keep]
            if lo.size:
                order = np.argsort(lo, kind="stable")
                lo, hi = lo[order], hi[order]
                reach = np.maximum.accumulate(hi)
                start = np.ones(lo.size, bool)
                start[1:] = lo[1:] > reach[:-1]  # touching intervals merge
                first = np.flatnonzero(start)
                lo = lo[first]
                hi = np.maximum.reduceat(hi, first)
        self.lo, self.hi = lo, hi

    # ---- constructors
    @staticmethod
    def empty():
        return IntervalSet()

    @staticmethod
    def openclosed(a, b):
        return IntervalSet([a], [b])

    @staticmethod
    def from_pairs(pairs):
        pairs = np.asarray(list(pairs), dtype=np.int64).reshape(-1, 2)
        return IntervalSet(pairs[:, 0], pairs[:, 1])

    # ---- queries
    def pairs(self):
        return list(zip(self.lo.tolist(), self.hi.tolist()))

    def is_empty(self):
        return self.lo.size == 0

    def length(self):
        """Number of integer points = accumulated length in frames (utils.p_len)."""
        return int((self.hi - self.lo).sum())

    def __len__(self):
        return int(self.lo.size)

    def __eq__(self, other):
        return isinstance(other, IntervalSet) and np.array_equal(self.lo, other.lo) and np.array_equal(self.hi, other.hi)

    def __repr__(self):
        return " | ".join(f"({a},{b}]" for a, b in self.pairs()) or "()"

    # ---- algebra: one sweep over the merged endpoints; state k after coordinate u_k holds on (u_k, u_{k+1}]
    def _sweep(self, other, keep):
        if self.is_empty() and other.is_empty():
            return IntervalSet()
        pts = np.concatenate([self.lo, self.hi, other.lo, other.hi])
        da = np.concatenate([np.ones(self.lo.size, np.int64), -np.ones(self.hi.size, np.int64),
                             np.zeros(other.lo.size + other.hi.size, np.int64)])
        db = np.concatenate([np.zeros(self.lo.size + self.hi.size, np.int64), np.ones(other.lo.size, np.int64),
                             -np.ones(other.hi.size, np.int64)])
        u, inv = np.unique(pts, return_inverse=True)
        in_a = np.cumsum(np.bincount(inv, weights=da, minlength=u.size)) > 0.5
        in_b = np.cumsum(np.bincount(inv, weights=db, minlength=u.size)) > 0.5
        on = keep(in_a, in_b)[:-1]
        return IntervalSet(u[:-1][on], u[1:][on])

    def __or__(self, other):
        return IntervalSet(np.concatenate([self.lo, other.lo]), np.concatenate([self.hi, other.hi]))

    def __and__(self, other):
        return self._sweep(other, lambda a, b: a & b)

    def __sub__(self, other):
        return self._sweep(other, lambda a, b: a & ~b)

    def contains(self, other):
        """True when every point of `other` lies in this set (portion's Interval.contains for an interval argument)."""
        return (other - self).is_empty()

    def contains_each(self, lo, hi):
        """contains(openclosed(lo[i], hi[i])) for many intervals at once: an interval lies in a union of disjoint, merged
        intervals iff one of them covers it (empty intervals are contained in everything)."""
        lo, hi = np.asarray(lo, np.int64), np.asarray(hi, np.int64)
        if self.is_empty():
            return hi <= lo
        i = np.searchsorted(self.lo, lo, side="right") - 1
        covered = (i >= 0) & (hi <= self.hi[np.maximum(i, 0)])
        return covered | (hi <= lo)

    def overlaps(self, other):
        return not (self & other).is_empty()
