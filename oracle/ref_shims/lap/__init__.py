"""Stand-in for lap 0.5.12 `lapjv` (un-vendored C++ Jonker-Volgenant).

Published behaviour restated: with `extend_cost=True` (or a finite `cost_limit`) lap solves the
square (n+m)x(n+m) problem whose top-left block is `cost`, whose off-diagonal blocks are filled
with `cost_limit/2` (or `cost.max()+1` when no limit) and whose bottom-right block is zero; rows
assigned to a dummy column are reported as -1. scipy's LSA on that explicit matrix gives the same
assignment whenever the optimum over real pairs is unique.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


def lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    if n == 0 or m == 0:
        x = -np.ones(n, dtype=int)
        y = -np.ones(m, dtype=int)
        return (0.0, x, y) if return_cost else (x, y)
    if n != m and not extend_cost and not cost_limit < np.inf:
        raise ValueError("Square cost array expected. Pass extend_cost=True.")
    if extend_cost or cost_limit < np.inf:
        N = n + m
        ext = np.empty((N, N), dtype=np.float64)
        ext[:] = cost_limit / 2.0 if cost_limit < np.inf else cost.max() + 1
        ext[n:, m:] = 0
        ext[:n, :m] = cost
    else:
        ext = cost
    r, c = linear_sum_assignment(ext)
    x = -np.ones(n, dtype=int)
    y = -np.ones(m, dtype=int)
    for i, j in zip(r, c):
        if i < n and j < m:
            x[i] = j
            y[j] = i
    opt = float(cost[x >= 0, x[x >= 0]].sum()) if (x >= 0).any() else 0.0
    return (opt, x, y) if return_cost else (x, y)
