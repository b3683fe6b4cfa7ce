#!/usr/bin/env python
"""Would an Ukkonen cut-off help the bit-sliced edit-distance rows kernel?

The kernel evaluates all L rows of the DP column for a warp-step = 32 lanes (independent text positions) x 32 gRNA.  A
cut-off can skip the rows below the last ACTIVE row - the deepest row i with D[i][j] <= K - but only if it is skipped for
all 32 x 32 cells of the step (the row loop is unrolled straight-line code over registers shared by the warp).  This
simulates the unit-cost DP on uniform random text and gRNA (numpy, CPU) and prints the distribution of the last active row
per cell and per warp-step, i.e. the fraction of row work a perfect warp-granular cut-off would save.

    python tools/ukkonen_hist.py [K] [L] > profiles/r02_ukkonen_active_rows_K5.json
"""
import json
import sys

import numpy as np


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rng = np.random.Generator(np.random.PCG64(1))
    n_steps, lanes, group, chars = 400, 32, 32, 64          # steps sampled; each lane walks `chars` characters after warm-up
    cell_hist = np.zeros(L + 1, dtype=np.int64)
    step_hist = np.zeros(L + 1, dtype=np.int64)
    for _ in range(n_steps):
        q = rng.integers(0, 4, size=(group, L))
        text = rng.integers(0, 4, size=(lanes, chars + L + K))
        # D[lane, gRNA, row]; row 0 = 0 (free start), column recurrence of the semi-global edit distance
        D = np.broadcast_to(np.arange(L + 1), (lanes, group, L + 1)).copy()
        for j in range(text.shape[1]):
            mis = (q[None, :, :] != text[:, j, None, None]).astype(np.int64)          # lanes x group x L
            new = np.empty_like(D)
            new[:, :, 0] = 0
            for i in range(1, L + 1):
                new[:, :, i] = np.minimum(np.minimum(D[:, :, i - 1] + mis[:, :, i - 1], D[:, :, i] + 1), new[:, :, i - 1] + 1)
            D = new
            if j >= L + K:
                active = D <= K                                                       # lanes x group x (L+1)
                last = (active * np.arange(L + 1)).max(axis=2)                        # deepest active row per cell
                np.add.at(cell_hist, last.ravel(), 1)
                step_hist[int(last.max())] += 1
    cells, steps = int(cell_hist.sum()), int(step_hist.sum())
    rows_needed_step = float((step_hist * np.minimum(np.arange(L + 1) + 1, L)).sum()) / steps     # rows up to last active + 1
    out = {"K": K, "L": L, "cells": cells, "warp_steps": steps,
           "last_active_row_per_cell": {str(i): cell_hist[i] / cells for i in range(L + 1) if cell_hist[i]},
           "last_active_row_per_warp_step": {str(i): step_hist[i] / steps for i in range(L + 1) if step_hist[i]},
           "mean_rows_a_warp_granular_cutoff_must_still_compute": rows_needed_step,
           "row_work_saved_by_a_perfect_warp_cutoff": 1.0 - rows_needed_step / L,
           "note": "uniform random text and gRNA; a cell's column must be computed one row past its last active row"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
