"""libstdc++-exact random numbers for reference-compatible initial data.

The reference seeds `std::mt19937 gen(4302529u)` (S6/mgrid_ntl.cpp:35-36) and draws every initial field with
`std::uniform_real_distribution<double>(-pi, pi)` (S6/modules_indiv.h:19).  One double consumes two 32-bit
outputs, low word first: value = ((lo + hi*2^32) / 2^64) * 2*pi - pi.  numpy's legacy RandomState uses the
same MT19937 seeding (init_genrand), so the stream is reproduced bit for bit.
"""
from __future__ import annotations

import numpy as np


class StdMT19937:
    def __init__(self, seed: int = 4302529):
        self._rs = np.random.RandomState(seed)

    def uniform_pm_pi(self, n: int) -> np.ndarray:
        raw = self._rs.randint(0, 2 ** 32, size=2 * n, dtype=np.uint64)
        lo = raw[0::2].astype(np.float64)
        hi = raw[1::2].astype(np.float64)
        canon = (lo + hi * 4294967296.0) / 18446744073709551616.0
        canon = np.minimum(canon, np.nextafter(1.0, 0.0))
        return canon * (np.pi - (-np.pi)) + (-np.pi)
