"""Restatement of scikit-learn 0.22.2 ``sklearn.utils.linear_assignment_``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Parity unpinned: the module
is a third-party dependency of the reference (pinned at
``/root/reference/environment.yml:18``, imported at
``/root/reference/tracking/sort/sort.py:26``, called at ``sort.py:206``) that is
neither vendored under ``/root/reference`` nor installable here.  This file
restates its published algorithm — the matrix form of Kuhn-Munkres that also
shipped as ``scipy/optimize/_hungarian.py`` in SciPy <= 1.3 — step for step,
because the tracker's results depend on *which* optimal assignment is returned
(zero-IoU ties with ``iou_threshold=0``).

Semantics that matter and are kept:
  * all arithmetic happens in the dtype of the input (float32 at the call site),
    including the in-place ``+= minval`` / ``-= minval`` pair of step 6;
  * the matrix is transposed when it has more rows than columns;
  * "first zero" searches are row-major;
  * the result is the list of starred cells sorted lexicographically.
"""
import numpy as np

_STAR = 1
_PRIME = 2


class _State:
    __slots__ = ("C", "flipped", "row_free", "col_free", "mark", "z0", "trail")

    def __init__(self, cost):
        cost = np.atleast_2d(cost)
        self.flipped = cost.shape[1] < cost.shape[0]
        self.C = np.array(cost.T if self.flipped else cost, copy=True)
        n, m = self.C.shape
        self.row_free = np.ones(n, dtype=bool)
        self.col_free = np.ones(m, dtype=bool)
        self.mark = np.zeros((n, m), dtype=np.int64)
        self.z0 = (0, 0)
        self.trail = np.zeros((n + m, 2), dtype=np.int64)

    def uncover_all(self):
        self.row_free[:] = True
        self.col_free[:] = True


def _reduce_and_star(st):
    # row reduction, then greedy starring of zeros in row-major order
    st.C -= st.C.min(axis=1)[:, None]
    rr, cc = np.nonzero(st.C == 0)
    for r, c in zip(rr, cc):
        if st.col_free[c] and st.row_free[r]:
            st.mark[r, c] = _STAR
            st.col_free[c] = False
            st.row_free[r] = False
    st.uncover_all()
    return _cover_starred_columns


def _cover_starred_columns(st):
    stars = st.mark == _STAR
    st.col_free[stars.any(axis=0)] = False
    if stars.sum() < st.C.shape[0]:
        return _prime_zeros
    return None


def _prime_zeros(st):
    zeros = (st.C == 0).astype(np.int64)
    open_zeros = zeros * st.row_free[:, None]
    open_zeros *= st.col_free.astype(np.int64)
    n, m = st.C.shape
    while True:
        r, c = np.unravel_index(np.argmax(open_zeros), (n, m))
        if open_zeros[r, c] == 0:
            return _shift_by_min
        st.mark[r, c] = _PRIME
        sc = int(np.argmax(st.mark[r] == _STAR))
        if st.mark[r, sc] != _STAR:
            st.z0 = (int(r), int(c))
            return _augment
        st.row_free[r] = False
        st.col_free[sc] = True
        open_zeros[:, sc] = zeros[:, sc] * st.row_free.astype(np.int64)
        open_zeros[r] = 0


def _augment(st):
    k = 0
    trail = st.trail
    trail[0] = st.z0
    while True:
        r = int(np.argmax(st.mark[:, trail[k, 1]] == _STAR))
        if st.mark[r, trail[k, 1]] != _STAR:
            break
        k += 1
        trail[k] = (r, trail[k - 1, 1])
        c = int(np.argmax(st.mark[trail[k, 0]] == _PRIME))
        if st.mark[r, c] != _PRIME:
            c = -1
        k += 1
        trail[k] = (trail[k - 1, 0], c)
    for i in range(k + 1):
        r, c = trail[i]
        st.mark[r, c] = 0 if st.mark[r, c] == _STAR else _STAR
    st.uncover_all()
    st.mark[st.mark == _PRIME] = 0
    return _cover_starred_columns


def _shift_by_min(st):
    if st.row_free.any() and st.col_free.any():
        col_min = np.min(st.C[st.row_free], axis=0)
        delta = np.min(col_min[st.col_free])
        st.C[~st.row_free] += delta
        st.C[:, st.col_free] -= delta
    return _prime_zeros


def hungarian(cost):
    """Starred cells (row, col) of the Munkres solution, row-major order."""
    st = _State(cost)
    step = None if 0 in np.shape(cost) else _reduce_and_star
    while step is not None:
        step = step(st)
    cells = np.array(np.nonzero(st.mark == _STAR)).T
    if st.flipped:
        cells = cells[:, ::-1]
    return cells


def linear_assignment(cost):
    """Drop-in for ``sklearn.utils.linear_assignment_.linear_assignment``."""
    pairs = hungarian(cost).tolist()
    pairs.sort()
    out = np.array(pairs, dtype=int)
    out.shape = (-1, 2)
    return out
