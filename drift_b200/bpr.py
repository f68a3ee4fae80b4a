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


def _factor_row(counts, cap, alpha, beta):
    """factor() for an int64 array of counts >= 1, bit for bit: true division of two ints, ``r ** beta``
    through the C library's pow (what CPython's float.__pow__ calls; NumPy's power differs in the last bit),
    then IEEE add / mul / div / max, which NumPy and CPython share."""
    import ctypes as C

    from . import _capi

    r = counts.astype(np.float64) / float(cap)
    pw = np.empty_like(r)
    _capi.check(_capi.load().drift_host_pow(r.ctypes.data_as(C.POINTER(C.c_double)), float(beta),
                                            pw.ctypes.data_as(C.POINTER(C.c_double)), len(r)))
    return np.maximum(0.1, 1 / (1 + alpha * pw))


def build_tables(link_cap, alpha: float, beta: float, guard: int = 4, max_count: int = 0):
    """Returns (table float64[], class_off int64[n_class+1], link_class int32[L]).

    Class c covers counts 0..K_c where K_c is ``guard`` counts past the first saturated count (the factor
    is exactly 0.1 from ``1 + alpha*r**beta >= 10`` on), or ``max_count`` (no link can hold more agents than
    exist) if that is smaller -- which also bounds the tables for parameters that saturate late or never
    (alpha = 0.01, beta = 1 saturates at count = 900 x capacity).  The device clamps larger counts to the
    last entry: exact in both cases.  Without ``max_count`` a row is cut 200 x capacity + 1000 counts out
    (then the clamp is an approximation beyond that count; Engine always passes ``max_count``).
    """
    link_cap = np.asarray(link_cap, dtype=np.int64)
    caps, link_class = np.unique(link_cap, return_inverse=True)
    tables, off = [], [0]
    for cap in caps.tolist():
        cap = max(int(cap), 1)
        hard = 200 * cap + 1000 if max_count <= 0 else int(max_count)
        # first saturated count, estimated (r_sat = (9/alpha)**(1/beta)); the loop below checks it on the real values
        est = hard
        if alpha > 0:
            est = int(min(float(hard), np.ceil(cap * (9.0 / alpha) ** (1.0 / beta)))) + guard + 2
        n = min(max(est, 8), hard)
        while True:
            row = _factor_row(np.arange(1, n + 1, dtype=np.int64), cap, alpha, beta)
            sat = np.nonzero(row == 0.1)[0]
            if len(sat) and sat[0] + guard <= n:
                row = row[:sat[0] + guard]
                break
            if n >= hard:
                break
            n = min(2 * n, hard)
        tables.append(np.concatenate([[1.0], row]))
        off.append(off[-1] + len(row) + 1)
    table = np.concatenate(tables) if tables else np.zeros(1)
    return np.ascontiguousarray(table, dtype=np.float64), np.array(off, dtype=np.int64), link_class.astype(np.int32)
