"""Stand-in for the third-party ``bandmat`` package on the reference's MLPG path.  TEST INFRASTRUCTURE ONLY.

``morgana/viz/synthesis.py:4-5`` imports ``bandmat`` (MattShannon/bandmat; **unpinned** in the reference's ``setup.py:12``,
not vendored, absent from this image and not installable -- no network).  The reference's ``MLPG`` uses five of its entry
points (``synthesis.py:33, 65, 72-74, 168``): ``band_c_bm``, ``zeros``, ``dot_mv_plus_equals``, ``dot_mm_plus_equals`` and
``linalg.solveh``.  This module restates the *published semantics* of exactly those on NumPy, so that the reference's own
``MLPG`` code runs unmodified:

* a ``BandMat(l, u, data, transposed)`` holds a square matrix with lower / upper bandwidths ``l`` / ``u`` in LAPACK's
  column-band layout: ``data[u + i - j, j] = A[i, j]`` (``.T`` shares ``data`` and flips ``transposed``);
* ``dot_mv_plus_equals(A, b, target)``: ``target += A @ b``; ``dot_mm_plus_equals(A, B, target_bm, diag)``:
  ``target_bm += A @ diag(d) @ B`` restricted to the target's band;
* ``linalg.solveh(A, b)``: solve with a symmetric positive-definite banded ``A`` by banded Cholesky -- here
  ``scipy.linalg.solveh_banded`` (LAPACK ``dpbsv``), the same factorisation ``bandmat.linalg.solveh`` performs
  (``cholesky`` + two banded triangular solves), fp64.

Because the third-party arithmetic itself is not available, MLPG's parity stays "pinned to the reference's code with a
named stand-in solver (scipy 1.18 ``solveh_banded``)", as SURVEY.md section 8c puts it -- not to bandmat's own bits.
"""
import sys
import types

import numpy as np
import scipy.linalg


class BandMat(object):
    def __init__(self, l, u, data, transposed=False):
        self.l, self.u, self.data, self.transposed = int(l), int(u), data, bool(transposed)
        assert self.data.shape[0] == self.l + self.u + 1

    @property
    def size(self):
        return self.data.shape[1]

    @property
    def T(self):
        return BandMat(self.u, self.l, self.data, not self.transposed)

    def diagonal(self, k):
        """``A[i, i + k]`` for the rows where it exists, as (first_row, values); ``-l <= k <= u``."""
        n = self.size
        lo, hi = max(0, -k), min(n, n - k)          # rows i with 0 <= i + k < n
        if hi <= lo:
            return lo, np.zeros((0,), dtype=self.data.dtype)
        if not self.transposed:
            return lo, self.data[self.u - k, lo + k:hi + k]       # data[u + i - j, j] with j = i + k
        return lo, self.data[self.l + k, lo:hi]                   # stored matrix is A^T with bandwidths (u, l)

    def add_to_diagonal(self, k, first_row, values):
        n = len(values)
        if n == 0:
            return
        if not self.transposed:
            self.data[self.u - k, first_row + k:first_row + k + n] += values
        else:
            self.data[self.l + k, first_row:first_row + n] += values

    def full(self):
        n = self.size
        out = np.zeros((n, n), dtype=self.data.dtype)
        for k in range(-self.l, self.u + 1):
            lo, v = self.diagonal(k)
            idx = np.arange(lo, lo + len(v))
            out[idx, idx + k] = v
        return out


def band_c_bm(l, u, mat_rect):
    return BandMat(l, u, np.asarray(mat_rect, dtype=np.float64))


def zeros(l, u, size):
    return BandMat(l, u, np.zeros((l + u + 1, size), dtype=np.float64))


def dot_mv_plus_equals(a_bm, b, target):
    for k in range(-a_bm.l, a_bm.u + 1):
        lo, v = a_bm.diagonal(k)
        if len(v):
            target[lo:lo + len(v)] += v * b[lo + k:lo + k + len(v)]


def dot_mm_plus_equals(a_bm, b_bm, target_bm, diag=None):
    n = a_bm.size
    assert b_bm.size == n and target_bm.size == n
    d = np.ones((n,)) if diag is None else np.asarray(diag, dtype=np.float64)
    for ka in range(-a_bm.l, a_bm.u + 1):
        lo_a, va = a_bm.diagonal(ka)                 # A[i, i + ka], i in [lo_a, lo_a + len)
        for kb in range(-b_bm.l, b_bm.u + 1):
            k = ka + kb
            if k < -target_bm.l or k > target_bm.u:
                continue
            lo_b, vb = b_bm.diagonal(kb)             # B[m, m + kb], m in [lo_b, lo_b + len)
            # C[i, i + k] += A[i, i + ka] d[i + ka] B[i + ka, i + ka + kb]: rows i with i in A's range and i + ka in B's
            first = max(lo_a, lo_b - ka)
            last = min(lo_a + len(va), lo_b + len(vb) - ka)
            if last <= first:
                continue
            i = np.arange(first, last)
            target_bm.add_to_diagonal(k, first, va[i - lo_a] * d[i + ka] * vb[i + ka - lo_b])


def solveh(a_bm, b):
    """Banded Cholesky solve of a symmetric positive-definite system (upper form of LAPACK's ``dpbsv``)."""
    u = a_bm.u
    ab = np.zeros((u + 1, a_bm.size), dtype=np.float64)
    for k in range(0, u + 1):
        lo, v = a_bm.diagonal(k)
        ab[u - k, lo + k:lo + k + len(v)] = v
    return scipy.linalg.solveh_banded(ab, np.asarray(b, dtype=np.float64), lower=False)


def install():
    """Register ``bandmat`` and ``bandmat.linalg`` in ``sys.modules`` (no-op if a real bandmat is importable)."""
    if 'bandmat' in sys.modules and getattr(sys.modules['bandmat'], '__morgana_b200_standin__', None) is None:
        return sys.modules['bandmat']
    bm = types.ModuleType('bandmat')
    bm.__morgana_b200_standin__ = 'numpy restatement of the five entry points MLPG uses; solver = scipy.linalg.solveh_banded'
    bm.BandMat, bm.band_c_bm, bm.zeros = BandMat, band_c_bm, zeros
    bm.dot_mv_plus_equals, bm.dot_mm_plus_equals = dot_mv_plus_equals, dot_mm_plus_equals
    bla = types.ModuleType('bandmat.linalg')
    bla.solveh = solveh
    bm.linalg = bla
    sys.modules['bandmat'], sys.modules['bandmat.linalg'] = bm, bla
    return bm
