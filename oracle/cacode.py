"""GPS L1 C/A Gold-code generator (oracle; test infrastructure only).

Restates ``SDR_MATLAB-main/acqtckpos/generateCAcode.m:16-64``: two 10-stage
LFSRs in +-1 arithmetic (G1 feedback taps 3 and 10, G2 feedback taps
2,3,6,8,9,10, both loaded with all -1), G2 circularly delayed by ``g2s(PRN)``
chips, output ``-(g1 .* g2)``.
"""
from __future__ import annotations

import numpy as np

# generateCAcode.m:16-24 -- G2 delay in chips per PRN (GPS 1..32, then SBAS).
G2_SHIFT = (
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251,
    252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
    473, 474, 509, 512, 513, 514, 515, 516, 859, 860,
    861, 862,
    145, 175, 52, 21, 237, 235, 886, 657,
    634, 762, 355, 1012, 176, 603, 130, 359, 595, 68,
    386,
)

# IS-GPS-200, Table 3-Ia: first ten C/A chips, octal, PRN 1..32.
IS_GPS_200_FIRST10_OCTAL = (
    0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454,
    0o1626, 0o1504, 0o1642, 0o1750, 0o1764, 0o1772, 0o1775, 0o1776,
    0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706,
    0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712,
)


def _lfsr(taps: tuple[int, ...]) -> np.ndarray:
    """1023 outputs of a 10-stage +-1 shift register (generateCAcode.m:32-42, 47-57)."""
    reg = -np.ones(10, dtype=np.float64)          # :34 / :49
    out = np.zeros(1023, dtype=np.float64)
    for i in range(1023):                         # :37 / :52
        out[i] = reg[9]                           # output = stage 10
        fb = 1.0
        for t in taps:                            # product of tapped stages
            fb *= reg[t - 1]
        reg[1:] = reg[:-1].copy()                 # shift
        reg[0] = fb
    return out


def generate_ca_code(prn: int) -> np.ndarray:
    """Return the 1x1023 +-1 C/A code of ``prn`` (1-based) as float64.

    Follows generateCAcode.m:27 (shift lookup), :61 (rotate G2 right by the
    shift) and :64 (``-(g1.*g2)``).
    """
    if not 1 <= prn <= len(G2_SHIFT):
        raise ValueError(f"PRN {prn} out of range 1..{len(G2_SHIFT)}")
    shift = G2_SHIFT[prn - 1]
    g1 = _lfsr((3, 10))
    g2 = _lfsr((2, 3, 6, 8, 9, 10))
    g2 = np.concatenate((g2[1023 - shift:], g2[:1023 - shift]))   # :61
    return -(g1 * g2)                                             # :64
