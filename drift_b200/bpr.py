"""Host-built BPR factor tables.

The reference computes ``max(0.1, 1 / (1 + alpha * (count / capacity) ** beta))``
(lib/managers/traffic_manager.py:88-107) with Python floats, i.e. glibc ``pow``; CUDA's ``pow``
is not bit-identical to it.  The factor only depends on (count, capacity) and saturates at exactly
0.1 once ``1 + alpha*r**beta >= 10``, so the host tabulates it per distinct capacity with the very
same Python expression and the device looks it up (a few hundred classes, < 1 M doubles).
"""
from __future__ import annotations

import numpy as np


def factor(count: int, capacity: int, alpha: float, beta: float) -> float:
    if count == 0:
        return 1.0
    r = count / capacity
    m = 1 + alpha * (r ** beta)
    return max(0.1, 1 / m)


def build_tables(link_cap, alpha: float, beta: float, guard: int = 4):
    """Returns (table float64[], class_off int64[n_class+1], link_class int32[L]).

    Class c covers counts 0..K_c where K_c is ``guard`` counts past the first saturated count; the
    device clamps larger counts to the last entry (0.1).
    """
    link_cap = np.asarray(link_cap, dtype=np.int64)
    caps, link_class = np.unique(link_cap, return_inverse=True)
    tables, off = [], [0]
    for cap in caps.tolist():
        cap = int(cap)
        row = [1.0]
        n = 1
        sat = 0
        while True:
            f = factor(n, cap, alpha, beta)
            row.append(f)
            sat = sat + 1 if f == 0.1 else 0
            if sat >= guard:
                break
            n += 1
            if n > 200 * max(cap, 1) + 1000:  # alpha ~ 0: never saturates; the tail keeps the last value
                break
        tables.append(np.array(row, dtype=np.float64))
        off.append(off[-1] + len(row))
    table = np.concatenate(tables) if tables else np.zeros(1)
    return table, np.array(off, dtype=np.int64), link_class.astype(np.int32)
