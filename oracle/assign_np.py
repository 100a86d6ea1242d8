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


def lsap_scipy_restated(cost):
    """scipy.optimize.linear_sum_assignment restated step by step (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp, the
    shortest-augmenting-path solver of Crouse that the StrongSORT plugins call, /root/reference/plugins/track/strong_sort/
    sort/linear_assignment.py:55, bpbreid_strong_sort/sort/linear_assignment.py:56) INCLUDING its tie-breaking, which decides
    the result whenever entries are equal — the clamped `max_distance + 1e-5` entries of min_cost_matching always are:
      * a tall matrix is transposed; rows are augmented in index order;
      * the not-yet-scanned columns live in a vector filled in REVERSE (nc-1 .. 0) from which the chosen column is removed
        by swapping in the last element;
      * per step the vector is walked in order; a column replaces the current choice when its path cost is lower, or equal
        and the column is unassigned — i.e. the result is the LAST unassigned column with the minimum cost if there is one,
        else the FIRST column with the minimum cost;
      * r = minVal + cost[i, j] - u[i] - v[j] evaluated left to right in float64.
    tests/test_oracle_cpu.py pins this restatement to scipy itself on tie-heavy matrices; the device solver
    (csrc/lsap_scipy.cuh) follows it operation by operation. Returns (row_ind, col_ind) like scipy."""
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    transpose = nc < nr
    C = cost.T.copy() if transpose else cost
    if transpose:
        nr, nc = nc, nr
    u, v = np.zeros(nr), np.zeros(nc)
    spc = np.empty(nc)
    path = np.full(nc, -1, dtype=np.int64)
    col4row = np.full(nr, -1, dtype=np.int64)
    row4col = np.full(nc, -1, dtype=np.int64)
    for cur in range(nr):
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        SR = np.zeros(nr, dtype=bool); SC = np.zeros(nc, dtype=bool)
        spc[:] = np.inf
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            index, lowest = -1, np.inf
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + C[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for i2 in range(nr):
            if SR[i2] and i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if SC[j2]:
                v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    if transpose:
        order = np.argsort(col4row, kind="stable")
        return col4row[order], order
    return np.arange(nr, dtype=np.int64), col4row.copy()
