"""Synthetic GPS L1 C/A IF generator (test infrastructure only).

The reference ships no recording (``initParameters.m:21`` points at a local
Windows path), so inputs are synthesised in the recordings' format: int8,
I then Q interleaved (``acquisition.m:36``), or int8 real, or int16 I/Q.

Sample ``n0`` (0-based, counted from the start of the virtual file) is

    x[n0] = sum_s A_s * b_s(floor(n0 / (20 N))) * CA_s[floor((n0 - tau_s) * fc/Fs) mod 1023]
                  * exp(-j (2 pi (IF + f_s) n0 / Fs + phi_s))  +  sigma (w_I + j w_Q)

The carrier sits at *negative* frequency because the reference wipes off with
``exp(+j...)`` (``acquisition.m:43,56``).  Noise comes from
``default_rng([seed, ms_index])`` per 1-ms block, so any window of a long
virtual file is reproducible.  Expected coarse result for an SV:
``Doppler`` = grid value nearest ``f_s``, ``codedelay = (N - 1 - tau_s) mod N``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from .cacode import generate_ca_code


@dataclass
class SatSpec:
    prn: int
    doppler_hz: float
    codedelay: int            # the value acquisition should report (0-based lag)
    amplitude: float = 0.4    # LSB
    phase: float = 0.0        # rad


@dataclass
class SynthSpec:
    fs: float = 58e6
    if_hz: float = 4.58e6
    code_hz: float = 1.023e6
    samples_per_ms: int = 58000
    sigma: float = 16.0       # noise std per component, LSB
    data_type: int = 2        # 1 real, 2 I/Q
    data_precision: int = 1   # 1 int8, 2 int16
    seed: int = 6102
    sats: List[SatSpec] = field(default_factory=list)


def _truth(prns, dopp, delays, amps) -> List[SatSpec]:
    return [SatSpec(p, f, d, a, 0.37 * i) for i, (p, f, d, a) in enumerate(zip(prns, dopp, delays, amps))]


# Truth tables seeded from the reference's saved results
# (Acquired_Opensky_5000.mat / nAcquired_Urban_5000.mat: fineFreq-IF and codedelay).
OPENSKY_TRUTH = _truth(
    (3, 4, 16, 22, 26, 27, 31, 32),
    (990.0, -3095.0, -305.0, 1565.0, 1835.0, -3225.0, 1045.0, 3345.0),
    (3683, 12701, 26051, 2610, 57908, 49778, 39064, 20170),
    (0.30, 0.27, 0.45, 0.33, 0.47, 0.38, 0.42, 0.37))
URBAN_TRUTH = _truth(
    (1, 3, 7, 11, 18, 22),
    (1200.0, 4285.0, 365.0, 405.0, -365.0, 3315.0),
    (22742, 1154, 10811, 24851, 15362, 2050),
    (2.2, 0.70, 0.50, 0.60, 0.48, 0.45))


def opensky_spec(seed: int = 6103, **kw) -> SynthSpec:
    return SynthSpec(sats=list(OPENSKY_TRUTH), seed=seed, **kw)


def urban_spec(seed: int = 6104, **kw) -> SynthSpec:
    return SynthSpec(fs=26e6, if_hz=0.0, samples_per_ms=26000, sats=list(URBAN_TRUTH), seed=seed, **kw)


def synth_samples(spec: SynthSpec, start_ms: int, n_ms: int) -> np.ndarray:
    """Complex float64 samples (before quantisation) of ms ``start_ms .. start_ms+n_ms-1``."""
    n = spec.samples_per_ms
    out = np.empty(n * n_ms, dtype=np.complex128)
    codes = {s.prn: generate_ca_code(s.prn) for s in spec.sats}
    for i in range(n_ms):
        ms = start_ms + i
        n0 = np.arange(ms * n, (ms + 1) * n, dtype=np.int64)
        x = np.zeros(n, dtype=np.complex128)
        for s in spec.sats:
            tau = (n - 1 - s.codedelay) % n
            chip = np.floor((n0 - tau) * (spec.code_hz / spec.fs)).astype(np.int64) % 1023
            bit_rng = np.random.default_rng([spec.seed, 7919, s.prn, ms // 20])
            bit = 1.0 if bit_rng.integers(0, 2) else -1.0
            f = spec.if_hz + s.doppler_hz
            if float(f).is_integer() and float(spec.fs).is_integer():
                # exact cycles: (f * n0) mod Fs in int64 (f*n0 < 2^63 for any 90 s window)
                cyc = ((int(f) * n0) % int(spec.fs)) / spec.fs
            else:
                cyc = (f * n0 / spec.fs) % 1.0
            ph = 2.0 * np.pi * cyc + s.phase
            x += (s.amplitude * bit) * codes[s.prn][chip] * np.exp(-1j * ph)
        rng = np.random.default_rng([spec.seed, ms])
        x += spec.sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        out[i * n:(i + 1) * n] = x
    return out


def synth_if(spec: SynthSpec, start_ms: int, n_ms: int) -> bytes:
    """The bytes a recording would hold for that window (little-endian, I first)."""
    x = synth_samples(spec, start_ms, n_ms)
    lo, hi, dt = (-128, 127, np.int8) if spec.data_precision == 1 else (-32768, 32767, "<i2")
    if spec.data_type == 2:
        iq = np.empty(2 * x.size, dtype=np.float64)
        iq[0::2] = x.real
        iq[1::2] = x.imag
    else:
        iq = x.real
    return np.clip(np.rint(iq), lo, hi).astype(dt).tobytes()


class VirtualFile:
    """A read-only, seekable stand-in for ``file.fid`` that synthesises bytes on demand."""

    def __init__(self, spec: SynthSpec):
        self.spec = spec
        self._pos = 0
        self._bpm = spec.samples_per_ms * spec.data_type * spec.data_precision   # bytes per ms

    def seek(self, offset: int, whence: int = 0) -> int:
        self._pos = offset if whence == 0 else self._pos + offset
        return self._pos

    def tell(self) -> int:
        return self._pos

    def read(self, nbytes: int) -> bytes:
        first = self._pos // self._bpm
        last = (self._pos + nbytes + self._bpm - 1) // self._bpm
        blob = synth_if(self.spec, first, last - first)
        off = self._pos - first * self._bpm
        self._pos += nbytes
        return blob[off:off + nbytes]
