"""Synthetic GPS L1 C/A IF recordings in the reference's file formats (product-side generator).

The reference ships no recording (``initParameters.m:21`` points at a local Windows path), so the
benchmarks and examples feed the library synthetic IF with known PRNs, Dopplers and code delays,
quantised like the Opensky / Urban recordings: int8 I then Q (``acquisition.m:36``), int8 real, or
int16 I/Q.  Spec: SURVEY.md Appendix C.  Uses the library's own C/A tables (``gnssacq_ca_code``);
``oracle/synth.py`` is an independent NumPy-only twin used by the tests, and
``tests/test_cabi.py::test_product_and_oracle_generators_agree`` keeps the two byte-identical.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import api


@dataclass
class Satellite:
    prn: int
    doppler_hz: float
    codedelay: int            # the 0-based lag acquisition reports: (N - 1 - true delay) mod N
    amplitude: float = 0.4    # LSB
    phase: float = 0.0


@dataclass
class Recording:
    fs: float = 58e6
    if_hz: float = 4.58e6
    code_hz: float = 1.023e6
    samples_per_ms: int = 58000
    sigma: float = 16.0
    data_type: int = 2        # 1 real, 2 I/Q
    data_precision: int = 1   # 1 int8, 2 int16
    seed: int = 6102
    sats: List[Satellite] = field(default_factory=list)

    def samples(self, start_ms: int, n_ms: int) -> np.ndarray:
        n = self.samples_per_ms
        out = np.empty(n * n_ms, dtype=np.complex128)
        codes = {s.prn: api.ca_code(s.prn).astype(np.float64) for s in self.sats}
        for i in range(n_ms):
            ms = start_ms + i
            n0 = np.arange(ms * n, (ms + 1) * n, dtype=np.int64)
            x = np.zeros(n, dtype=np.complex128)
            for s in self.sats:
                tau = (n - 1 - s.codedelay) % n
                chip = np.floor((n0 - tau) * (self.code_hz / self.fs)).astype(np.int64) % 1023
                bit = 1.0 if np.random.default_rng([self.seed, 7919, s.prn, ms // 20]).integers(0, 2) else -1.0
                f = self.if_hz + s.doppler_hz
                if float(f).is_integer() and float(self.fs).is_integer():
                    cyc = ((int(f) * n0) % int(self.fs)) / self.fs
                else:
                    cyc = (f * n0 / self.fs) % 1.0
                x += (s.amplitude * bit) * codes[s.prn][chip] * np.exp(-1j * (2.0 * np.pi * cyc + s.phase))
            rng = np.random.default_rng([self.seed, ms])
            x += self.sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
            out[i * n:(i + 1) * n] = x
        return out

    def read(self, start_ms: int, n_ms: int) -> bytes:
        """The bytes the recording holds for ms ``start_ms .. start_ms+n_ms-1`` (little-endian, I first)."""
        x = self.samples(start_ms, n_ms)
        lo, hi, dt = (-128, 127, np.int8) if self.data_precision == 1 else (-32768, 32767, "<i2")
        if self.data_type == 2:
            iq = np.empty(2 * x.size, dtype=np.float64)
            iq[0::2] = x.real
            iq[1::2] = x.imag
        else:
            iq = x.real
        return np.clip(np.rint(iq), lo, hi).astype(dt).tobytes()


def _sats(prns, dopp, delays, amps):
    return [Satellite(p, f, d, a, 0.37 * i) for i, (p, f, d, a) in enumerate(zip(prns, dopp, delays, amps))]


def opensky_recording(seed: int = 6103) -> Recording:
    """Opensky-shaped (58 MHz, IF 4.58 MHz, int8 I/Q); truth table from the reference's Acquired_Opensky_5000.mat."""
    return Recording(seed=seed, sats=_sats(
        (3, 4, 16, 22, 26, 27, 31, 32), (990.0, -3095.0, -305.0, 1565.0, 1835.0, -3225.0, 1045.0, 3345.0),
        (3683, 12701, 26051, 2610, 57908, 49778, 39064, 20170), (0.30, 0.27, 0.45, 0.33, 0.47, 0.38, 0.42, 0.37)))


def urban_recording(seed: int = 6104) -> Recording:
    """Urban-shaped (26 MHz, IF 0, int8 I/Q); truth table from the reference's nAcquired_Urban_5000.mat."""
    return Recording(fs=26e6, if_hz=0.0, samples_per_ms=26000, seed=seed, sats=_sats(
        (1, 3, 7, 11, 18, 22), (1200.0, 4285.0, 365.0, 405.0, -365.0, 3315.0),
        (22742, 1154, 10811, 24851, 15362, 2050), (2.2, 0.70, 0.50, 0.60, 0.48, 0.45)))
