"""ORACLE — threshold incomplete Cholesky ICT(p, tau), the CPU restatement behind ``dp_icholt_host`` (test infrastructure).

Stands in for ``ilupp.icholt(A, add_fill_in, threshold)`` (``uibk/deep_preconditioning/test.py:86``). ``ilupp`` 1.0.2
(``uv.lock:952``) is absent from this image and its source is not vendored in the reference: the VALUES are
**parity unpinned**; what is restated is the published scheme (Saad, *Iterative Methods for Sparse Linear Systems*,
§10.4, applied row-wise to the Cholesky factor) with the two dropping rules the ``ilupp`` arguments name:

* rule 1 (``threshold``): ``|L_ij| < threshold * ||A_i,0:i||_2`` -> dropped as soon as it is computed;
* rule 2 (``add_fill_in``): a row keeps at most ``nnz(A_i,0:i-1) + add_fill_in`` off-diagonal entries, the largest.

Plain Python loops: small cases only. The operations and their order are those of ``csrc/icholt.cu`` (bit-identical).
"""

from __future__ import annotations

import heapq
import math

import numpy as np


def icholt(rowptr, col, val, fill_in: int = 1, threshold: float = 0.1):
    """``(rowptr int32, col int32, val float64)`` of ``L`` from the CSR of ``tril(A)`` (sorted rows, diagonal last)."""
    n = len(rowptr) - 1
    rows_of_col = [[] for _ in range(n)]
    out_cols, out_vals, out_ptr = [], [], [0]
    for i in range(n):
        rs, re = int(rowptr[i]), int(rowptr[i + 1])
        assert re > rs and col[re - 1] == i, "diagonal must be stored, last in its row"
        acc = {int(col[p]): float(val[p]) for p in range(rs, re)}
        norm2 = 0.0
        for p in range(rs, re):
            norm2 += float(val[p]) * float(val[p])
        tau = threshold * math.sqrt(norm2)
        keep = (re - rs - 1) + fill_in
        heap = [j for j in acc if j < i]
        heapq.heapify(heap)
        queued = set(heap)
        cols, vals = [], []
        while heap:
            j = heapq.heappop(heap)
            s = acc.get(j, 0.0)
            js, je = out_ptr[j], out_ptr[j + 1] - 1
            a, b = 0, js
            while a < len(cols) and b < je:
                ca, cb = cols[a], out_cols[b]
                if ca == cb:
                    s -= vals[a] * out_vals[b]
                    a, b = a + 1, b + 1
                elif ca < cb:
                    a += 1
                else:
                    b += 1
            lij = s / out_vals[je]
            if abs(lij) < tau or lij == 0.0:
                continue
            cols.append(j), vals.append(lij)
            for r in rows_of_col[j]:
                if r < i and r not in queued:
                    queued.add(r)
                    heapq.heappush(heap, r)
        if len(cols) > keep:
            order = sorted(range(len(cols)), key=lambda q: -abs(vals[q]))[:keep]  # stable: ties keep the smaller column
            order.sort()
            cols, vals = [cols[q] for q in order], [vals[q] for q in order]
        d = acc[i]
        for v in vals:
            d -= v * v
        assert d > 0.0, "non-positive pivot"
        for c, v in zip(cols, vals):
            out_cols.append(c), out_vals.append(v)
            rows_of_col[c].append(i)
        out_cols.append(i), out_vals.append(math.sqrt(d))
        out_ptr.append(len(out_cols))
    return np.asarray(out_ptr, np.int32), np.asarray(out_cols, np.int32), np.asarray(out_vals, np.float64)
