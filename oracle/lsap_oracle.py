"""TEST INFRASTRUCTURE (oracle): pure-Python restatement of scipy.optimize.linear_sum_assignment
(scipy 1.18.1, the dependency the reference calls at src/utils/matcher.py:111,188; the algorithm is the
shortest-augmenting-path method of Crouse, "On implementing 2D rectangular assignment algorithms", 2016, as
implemented in scipy/optimize/rectangular_lsap/rectangular_lsap.cpp).  scipy ships only the compiled module
here, so this restatement is pinned empirically: tests/test_lsap_oracle.py checks it against scipy itself on
thousands of random, rectangular and tie-heavy integer matrices (identical assignments, not just equal cost).
The CUDA kernel csrc/lsap.cu follows this file step by step, including the order of the `remaining` list and
the tie rule, which is what makes its assignments bit-identical to scipy's.
"""
import math


def linear_sum_assignment(cost):
    """cost: list of rows (nr x nc), finite floats.  Returns (row_ind, col_ind) like scipy."""
    nr = len(cost)
    nc = len(cost[0]) if nr else 0
    if nr == 0 or nc == 0:
        return [], []
    transpose = nc < nr
    if transpose:
        cost = [[cost[i][j] for i in range(nr)] for j in range(nc)]
        nr, nc = nc, nr
    u = [0.0] * nr
    v = [0.0] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    for cur in range(nr):
        # ---- shortest augmenting path from row `cur` ----
        min_val = 0.0
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        SR = [False] * nr
        SC = [False] * nc
        spc = [math.inf] * nc
        sink = -1
        i = cur
        while sink == -1:
            index = -1
            lowest = math.inf
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + cost[i][j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == math.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        # ---- dual update ----
        u[cur] += min_val
        for i in range(nr):
            if SR[i] and i != cur:
                u[i] += min_val - spc[col4row[i]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        # ---- augment ----
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transpose:
        order = sorted(range(nr), key=lambda r: col4row[r])
        return [col4row[r] for r in order], order
    return list(range(nr)), col4row
