"""``TckResultCT, CN0_Eph, countinx = trackingCT(file, signal, track, Acquired)`` -- first (1 ms) stage.

Python twin of ``SDR_MATLAB-main/acqtckpos/trackingCT.m:1-212``: for every acquired satellite the conventional
DLL/PLL loop over ``track.msToProcessCT_1ms`` integration periods, the C/N0 estimate every 20 periods
(``:121-133``) and the navigation-bit edge index (``:176-212``).  File I/O stays on the host side of the boundary
(one ``seek`` + ``read`` of the segment the loops will walk through, instead of the reference's ``fread`` per
millisecond); the loops themselves run in ``libgnssacq.so`` on the GPU (``gnssacq_track``: every channel in
parallel, no host round trip per millisecond).  The reference's later stages (re-run from the bit edge, 10 ms
integrations, ``:214-533``) are not rebuilt.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Tuple

import numpy as np

from . import api
from .acquisition import config_from_structs, get_searcher

RECORD_FIELDS = ("P_i", "P_q", "E_i", "E_q", "L_i", "L_q", "pll_discri", "dll_discri", "rem_chip", "code_hz",
                 "carrier_hz", "rem_phase", "sample_end", "num_samples")


def trackParameters():
    """``track.*`` of initParameters.m:58-70 (the fields the conventional loop reads)."""
    return SimpleNamespace(CorrelatorSpacing=0.5, DLLBW=2.0, DLLDamp=0.707, DLLGain=0.1, PLLBW=15.0, PLLDamp=0.707,
                           PLLGain=0.25, msToProcessCT_1ms=1000, pdi=1)


def cn0_estimates(p_i: np.ndarray, p_q: np.ndarray, ms: float = 1e-3, pdi: int = 1, K: int = 20) -> np.ndarray:
    """trackingCT.m:121-133: moment-method C/N0 from every K prompt powers (``var`` = MATLAB's N-1 variance)."""
    zk = np.asarray(p_i, dtype=np.float64) ** 2 + np.asarray(p_q, dtype=np.float64) ** 2
    out = []
    for b in range(len(zk) // K):
        z = zk[b * K:(b + 1) * K]
        mean_zk, var_zk = z.mean(), z.var(ddof=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            na2 = np.sqrt(mean_zk ** 2 - var_zk)                                  # NaN when the variance dominates (:126)
            var_iq = 0.5 * (mean_zk - na2)
            out.append(abs(10.0 * np.log10(1.0 / (1.0 * ms * pdi) * na2 / (2.0 * var_iq))))
    return np.array(out)


def bit_edge_index(p_i: np.ndarray) -> int:
    """trackingCT.m:176-212: first i >= 600 (1-based) whose six predecessors have the other sign and whose 17
    successors have the same sign as P_i(i); returns ``mod(i, 20) - 1`` (0 when no such i exists, like the
    zero-initialised ``countinx``)."""
    s = np.sign(np.asarray(p_i, dtype=np.float64))
    n = len(s)
    for i in range(max(7, 600), n):                              # 1-based i = 7 .. length-1 (loop bound :179), i >= 600
        j = i - 1                                                # 0-based
        if j + 17 >= n:
            break
        if all(s[j - d] != s[j] for d in range(1, 7)) and all(s[j + d] == s[j] for d in range(1, 18)):
            return i % 20 - 1
    return 0


def trackingCT(file, signal, track, Acquired) -> Tuple[Dict[int, Dict[str, np.ndarray]], np.ndarray, np.ndarray]:
    """Stage 1 of trackingCT.m.  ``TckResultCT[prn]`` holds the per-period arrays of ``:153-172`` (P_i, P_q, E_i,
    E_q, L_i, L_q, PLLdiscri, DLLdiscri, codedelay, remChip, codeFreq, carrierFreq, remPhase, numSample, delayValue,
    absoluteSample, codedelay2); ``CN0_Eph`` is (periods // 20, n_sv); ``countinx`` the bit-edge index per SV.

    Deliberate deviation: ``codedelay[i]`` = AcqCodeDelay + the cumulative sum of THIS satellite's delayValue.
    trackingCT.m:161 writes ``sum(delayValue(1:Index))`` with a LINEAR index into the (n_sv x n_ms) matrix, which
    mixes the satellites' corrections and equals the per-satellite sum only when one SV is tracked."""
    sv = [int(p) for p in np.atleast_1d(Acquired["sv"])]
    n_ms = int(track.msToProcessCT_1ms)
    if not sv:
        return {}, np.zeros((0, 0)), np.zeros(0)
    N = int(signal.Sample)
    bps = int(file.dataType) * int(file.dataPrecision)
    code_delay = [int(c) for c in np.atleast_1d(Acquired["codedelay"])]
    fine = [float(f) for f in np.atleast_1d(Acquired["fineFreq"])]
    # trackingCT.m:60: every channel starts at sample (Sample - AcqCodeDelay + 1 + skip*Sample); read once what the
    # slowest/fastest code clock can reach in n_ms periods (+2 ms of slack), relative to skip*Sample
    first = int(file.skip) * N
    n_samples = (n_ms + 3) * N
    file.fid.seek(first * bps, 0)
    seg = file.fid.read(n_samples * bps)
    acq_defaults = SimpleNamespace(freqMin=-10000.0, freqStep=500.0, freqNum=41, datalen=1)
    s = get_searcher(config_from_structs(file, signal, acq_defaults, prns=[1]))
    s.track_load(seg)
    start = [api.Channel(prn=p, num_samples=0, sample_offset=N - cd + 1, carrier_hz=f, rem_phase=0.0,
                         code_hz=float(signal.codeFreqBasis), rem_chip=0.0) for p, cd, f in zip(sv, code_delay, fine)]
    loops = api.LoopParams(dll_bw=track.DLLBW, dll_damp=track.DLLDamp, dll_gain=track.DLLGain, pll_bw=track.PLLBW,
                           pll_damp=track.PLLDamp, pll_gain=track.PLLGain, spacing_chips=track.CorrelatorSpacing)
    recs = s.track(start, n_ms, loops)
    result: Dict[int, Dict[str, np.ndarray]] = {}
    cn0: List[np.ndarray] = []
    countinx = np.zeros(len(sv))
    for c, (p, cd) in enumerate(zip(sv, code_delay)):
        cols = {f: np.array([getattr(r, f) for r in recs[c]], dtype=np.float64) for f in RECORD_FIELDS}
        delay = cols["num_samples"] - N * int(track.pdi)                          # :80
        absolute = (cols["sample_end"] + first) * bps                             # ftell (:171), bytes
        result[p] = {
            "P_i": cols["P_i"], "P_q": cols["P_q"], "E_i": cols["E_i"], "E_q": cols["E_q"], "L_i": cols["L_i"],
            "L_q": cols["L_q"], "PLLdiscri": cols["pll_discri"], "DLLdiscri": cols["dll_discri"],
            "codedelay": cd + np.cumsum(delay), "remChip": cols["rem_chip"], "codeFreq": cols["code_hz"],
            "carrierFreq": cols["carrier_hz"], "remPhase": cols["rem_phase"], "numSample": cols["num_samples"],
            "delayValue": delay, "absoluteSample": absolute,
            "codedelay2": np.mod(absolute / bps, float(signal.Fs) * float(signal.ms)),   # :172
        }
        cn0.append(cn0_estimates(cols["P_i"], cols["P_q"], float(signal.ms), int(track.pdi)))
        countinx[c] = bit_edge_index(cols["P_i"])
    return result, np.stack(cn0, axis=1), countinx
