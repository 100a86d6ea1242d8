"""Linear-assignment oracle (test infrastructure).

The reference calls two third-party solvers:
  * ``lap.lapjv(cost, extend_cost=True, cost_limit=L)`` — /root/reference/plugins/track/byte_track/matching.py:37-48,
    /root/reference/plugins/track/oc_sort/association.py:187-191 (lap 0.5.12, un-vendored);
  * ``scipy.optimize.linear_sum_assignment`` — /root/reference/plugins/track/strong_sort/sort/linear_assignment.py:55.

lap's published extension: an (n+m)^2 matrix, real costs top-left, ``L/2`` (or ``max+1`` without a
limit) in the off-diagonal blocks, zeros bottom-right. Minimising it is the same as minimising
``sum(c_ij - L)`` over partial matchings, i.e. a rectangular assignment on ``min(c - L, 0)`` whose
zero-cost pairs are dropped. Both forms are provided; tests check they agree.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


def lapjv_extended(cost, cost_limit=np.inf):
    """x[i] = column of row i or -1, y[j] = row of column j or -1 (lap.lapjv semantics)."""
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    x = -np.ones(n, dtype=np.int64)
    y = -np.ones(m, dtype=np.int64)
    if n == 0 or m == 0:
        return x, y
    big = np.empty((n + m, n + m), dtype=np.float64)
    big[:] = cost_limit / 2.0 if cost_limit < np.inf else cost.max() + 1
    big[n:, m:] = 0.0
    big[:n, :m] = cost
    rows, cols = linear_sum_assignment(big)
    for i, j in zip(rows, cols):
        if i < n and j < m:
            x[i] = j
            y[j] = i
    return x, y


def partial_assignment(cost, cost_limit=np.inf):
    """Equivalent rectangular form (what the CUDA solver implements)."""
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    x = -np.ones(n, dtype=np.int64)
    y = -np.ones(m, dtype=np.int64)
    if n == 0 or m == 0:
        return x, y
    if cost_limit < np.inf:
        red = np.minimum(cost - cost_limit, 0.0)
    else:
        red = cost
    rows, cols = linear_sum_assignment(red)
    for i, j in zip(rows, cols):
        if cost_limit < np.inf and not red[i, j] < 0.0:
            continue
        x[i] = j
        y[j] = i
    return x, y
