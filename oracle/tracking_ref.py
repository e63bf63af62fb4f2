"""CPU oracle for the tracking correlators -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy float64 restatement of the data-parallel part of the reference's conventional tracking loop,
``acqtckpos/trackingCT.m:75-118`` (one integration period of one channel: code replicas at the tap offsets,
carrier replica, I/Q split, the six sums), generalised to any tap list so that it also covers the 25-tap bank of
``trackingCT_POS_updated_multicorrelator.m:41,207-260``; plus the scalar loop closure of ``trackingCT.m:24-66,
121-150`` and ``calcLoopCoef.m:41-45`` that a test needs to generate realistic channel states.

PARITY UNPINNED (no MATLAB/Octave here).  One known representational difference: MATLAB's colon operator builds
``a:d:b`` symmetrically from both ends, NumPy's ``a + d*arange(n)`` from the left; the two differ by at most one
ulp, which matters only when a code-phase sample lands within an ulp of a chip boundary before ``ceil``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from .cacode import generate_ca_code

CODE_LENGTH = 1023            # signal.codelength (initParameters.m:47)
CODE_FREQ_BASIS = 1.023e6     # signal.codeFreqBasis (initParameters.m:44)


def samples_of(raw_bytes, data_type: int, precision: int) -> np.ndarray:
    """trackingCT.m:85-95: int8 real / int8 I,Q interleaved / int16 I,Q with per-integration DC removal."""
    if precision == 2:
        s = np.frombuffer(raw_bytes, dtype="<i2").astype(np.float64)
        i, q = s[0::2], s[1::2]
        return (i - i.mean()) + 1j * (q - q.mean())
    s = np.frombuffer(raw_bytes, dtype=np.int8).astype(np.float64)
    if data_type == 2:
        return s[0::2] + 1j * s[1::2]
    return s.astype(np.complex128)


def correlate(x: np.ndarray, fs: float, prn: int, carrier_hz: float, rem_phase: float, code_hz: float,
              rem_chip: float, spacing: Sequence[float]):
    """One integration of one channel.  x: the numSample complex samples read at trackingCT.m:85-95.
    Returns (I[tap], Q[tap]) with I = sum(code .* imag(x .* carr)), Q = sum(code .* real(x .* carr))
    (trackingCT.m:112-117: the reference calls the imaginary part "Inphase")."""
    n = len(x)
    ca = generate_ca_code(prn)
    code = np.concatenate(([ca[-1]], ca, [ca[0]]))                     # :66  [Code(end) Code Code(1)]
    step = code_hz / fs
    k = np.arange(n, dtype=np.float64)
    carr_time = k / fs                                                  # :103 (0:numSample)./Fs, first numSample
    wave = (2.0 * np.pi * (carrier_hz * carr_time)) + rem_phase         # :104
    carrsig = np.exp(1j * wave)                                         # :106
    mixed = x * carrsig
    inphase, quadrature = mixed.imag, mixed.real                        # :112-113
    out_i, out_q = [], []
    for sp in spacing:
        t = (0.0 + sp + rem_chip) + step * k                            # :96-98 (left-built colon, see header)
        idx = np.ceil(t).astype(np.int64)                               # :99-101  Code(ceil(t)+1), 1-based
        # general form of the 1025-entry table: entry ceil(t) (0-based) = CA[(ceil(t)-1) mod 1023]
        chips = np.where((idx >= 0) & (idx <= CODE_LENGTH + 1), code[np.clip(idx, 0, CODE_LENGTH + 1)],
                         ca[(idx - 1) % CODE_LENGTH])
        out_i.append(float(np.sum(chips * inphase)))                    # :115-117
        out_q.append(float(np.sum(chips * quadrature)))
    return np.array(out_i), np.array(out_q)


def next_rem_chip(n: int, code_hz: float, fs: float, rem_chip: float, pdi: int = 1) -> float:
    """trackingCT.m:102: remChip = (t_CodePrompt(numSample) + step) - codeFreqBasis*ms*pdi."""
    step = code_hz / fs
    t_last = (0.0 + 0.0 + rem_chip) + step * float(n - 1)
    return (t_last + step) - CODE_FREQ_BASIS * 1e-3 * pdi


def next_rem_phase(n: int, carrier_hz: float, fs: float, rem_phase: float) -> float:
    """trackingCT.m:104-105: rem(Wave(numSample+1), 2*pi) (MATLAB rem: sign of the dividend)."""
    wave_end = (2.0 * np.pi * (carrier_hz * (float(n) / fs))) + rem_phase
    return float(np.fmod(wave_end, 2.0 * np.pi))


def num_samples(code_hz: float, fs: float, rem_chip: float, pdi: int = 1) -> int:
    """trackingCT.m:78: round((codelength*pdi - remChip)/(codeFreq/Fs)) (MATLAB round: half away from zero)."""
    v = (CODE_LENGTH * pdi - rem_chip) / (code_hz / fs)
    return int(np.floor(v + 0.5)) if v >= 0 else -int(np.floor(-v + 0.5))


def calc_loop_coef(lbw: float, zeta: float, k: float):
    """calcLoopCoef.m:41-45."""
    wn = lbw * 8 * zeta / (4 * zeta ** 2 + 1)
    return k / (wn * wn), 2.0 * zeta / wn


@dataclass
class ChannelState:
    """The scalars trackingCT.m carries from one integration to the next (lines 42-58)."""
    prn: int
    carrier_basis_hz: float
    carrier_hz: float
    code_hz: float = CODE_FREQ_BASIS
    rem_chip: float = 0.0
    rem_phase: float = 0.0
    sample_pos: int = 0                 # file position in samples (ftell / (precision*type))
    code_out_last: float = 0.0
    dll_last: float = 0.0
    carr_out_last: float = 0.0
    pll_last: float = 0.0
    history: List[dict] = field(default_factory=list)


def close_loops(st: ChannelState, i_taps, q_taps, n: int, fs: float, *, dll=(2.0, 0.707, 0.1), pll=(15.0, 0.707, 0.25)):
    """trackingCT.m:135-150 for taps ordered [early, prompt, late]; advances the state by one integration."""
    tau1c, tau2c = calc_loop_coef(*dll)
    tau1p, tau2p = calc_loop_coef(*pll)
    e = np.hypot(i_taps[0], q_taps[0])
    late = np.hypot(i_taps[2], q_taps[2])
    dll_d = 0.5 * (e - late) / (e + late)
    code_out = st.code_out_last + (tau2c / tau1c) * (dll_d - st.dll_last) + dll_d * (0.001 / tau1c)
    pll_d = np.arctan(q_taps[1] / i_taps[1]) / (2 * np.pi)
    carr_out = st.carr_out_last + (tau2p / tau1p) * (pll_d - st.pll_last) + pll_d * (0.001 / tau1p)
    st.rem_chip = next_rem_chip(n, st.code_hz, fs, st.rem_chip)
    st.rem_phase = next_rem_phase(n, st.carrier_hz, fs, st.rem_phase)
    st.sample_pos += n
    st.dll_last, st.code_out_last = dll_d, code_out
    st.pll_last, st.carr_out_last = pll_d, carr_out
    st.code_hz = CODE_FREQ_BASIS - code_out
    st.carrier_hz = st.carrier_basis_hz + carr_out
    st.history.append(dict(P_i=i_taps[1], P_q=q_taps[1], dll=dll_d, pll=pll_d, code_hz=st.code_hz,
                           carrier_hz=st.carrier_hz, n=n))
    return st
