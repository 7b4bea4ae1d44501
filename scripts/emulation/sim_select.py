"""CPU emulation of selection_passes (csrc/pn2_grouping.cu): the reference's swap-based selection sort
reproduced WITHOUT the permuted array -- a table of displaced elements + one bit per position -- must give
the reference's outputs (first k AND the tail of the permutation) bit for bit, ties and NaNs included."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def select_row(row, k):
    n = len(row)
    bits = np.zeros(n, bool)
    table = {}  # pos -> (value, orig idx)
    out_v, out_i = np.empty(n, np.float32), np.empty(n, np.int32)
    passes = min(k, n)
    for s in range(passes):
        cs = table[s] if bits[s] else (row[s], s)
        best = None  # (value, pos, idx)
        for t in range(s + 1, n):
            if bits[t]:
                continue
            d = row[t]
            if d == d and (best is None or d < best[0]):
                best = (d, t, t)
        for tp, (d, oi) in table.items():
            if tp > s and d == d and (best is None or d < best[0] or (d == best[0] and tp < best[1])):
                best = (d, tp, oi)
        move = best is not None and best[0] < cs[0]
        out_v[s], out_i[s] = (best[0], best[2]) if move else cs
        if move:
            table[best[1]] = cs
            bits[best[1]] = True
    for t in range(passes, n):
        out_v[t], out_i[t] = table[t] if bits[t] else (row[t], t)
    return out_i, out_v


def main():
    rs = np.random.RandomState(0)
    cases = [rs.random_sample((1, 6, 50)), rs.randint(0, 4, (1, 8, 60)), rs.randint(0, 2, (1, 5, 40)),
             np.zeros((1, 3, 33)), rs.randint(0, 3, (1, 6, 9))]
    nan = rs.randint(0, 4, (1, 6, 40)).astype(np.float32)
    nan[0, :, ::7] = np.nan
    nan[0, 2, 0] = np.nan
    inf = rs.randint(0, 3, (1, 4, 30)).astype(np.float32)
    inf[inf == 2] = np.inf
    cases += [nan, inf]
    for d in cases:
        d = np.ascontiguousarray(d, np.float32)
        for k in (1, 3, 8, d.shape[2], d.shape[2] + 5):
            ei, eo = orc.select_top_k(k, d)
            for r in range(d.shape[1]):
                gi, go = select_row(d[0, r], k)
                assert np.array_equal(gi, ei[0, r]), (k, r, gi, ei[0, r])
                assert np.array_equal(go.view(np.uint32), eo[0, r].view(np.uint32)), (k, r)
    print("sim_select OK: %d cases" % len(cases))


if __name__ == "__main__":
    main()
