"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device's counter-based departure draw.

The reference draws ``random.random()`` (MT19937) per waiting agent per tick (lib/agent.py:180); a
sequential stream cannot be consumed in parallel, so the device-driven demand mode uses
Philox4x32-10 (Salmon et al., SC'11): one block = 128 bits = the draws of two agents, counter
(global agent id >> 1, tick), key = seed; the even agent takes words 0-1, the odd one words 2-3,
and the double is built the way MT's genrand_res53 does.  This file pins that generator
independently of the CUDA code.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def uniform53(seed, agent, tick):
    """Vectorised over ``agent`` (uint32 array); ``tick`` scalar.  float64 in [0, 1)."""
    agent = np.asarray(agent, dtype=np.uint64)
    odd = (agent & np.uint64(1)).astype(bool)
    c0 = (agent >> np.uint64(1)) & MASK
    c1 = np.full_like(c0, np.uint64(tick))
    c2 = np.full_like(c0, np.uint64(0x9E3779B9))
    c3 = np.full_like(c0, np.uint64(0xBB67AE85))
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    a = (np.where(odd, c2, c0) >> np.uint64(5)).astype(np.float64)
    b = (np.where(odd, c3, c1) >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)
