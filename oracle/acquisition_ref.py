"""NumPy float64 restatement of ``acquisition.m`` (oracle; test infrastructure only).

Every function cites the lines of
``/root/reference/SDR_MATLAB-main/acqtckpos/acquisition.m`` it follows.  Index
conventions: MATLAB is 1-based; everything returned here that the reference
stores 0-based (``codedelay = codePhase-1``) is 0-based, everything else says
which it is.

PARITY UNPINNED (no MATLAB/Octave here, no input recording in the reference):
see ``oracle/__init__.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .cacode import generate_ca_code


@dataclass
class CoarseRow:
    """One PRN's coarse-search outcome (acquisition.m:62-74), acquired or not."""
    prn: int
    acquired: bool
    code_phase: int        # codePhase-1 (0-based lag), what Acquired.codedelay stores (:74)
    doppler_bin: int       # fbin-1 (0-based)
    doppler_hz: float      # :64
    peak: float            # :63
    noise_meansq: float    # denominator of :67-68
    snr_db: float          # :67
    runner_up: float = float("nan")   # best cell outside the winner (diagnostic for tie tolerance)


# --------------------------------------------------------------------------- IF ingest
def read_if_block(file, signal, n_ms: int) -> np.ndarray:
    """acquisition.m:27-38 -- absolute seek, read ``n_ms`` ms, de-interleave.

    ``file.fid`` is any binary file-like object (``seek``/``read``).  A short
    read returns fewer samples, as MATLAB's ``fread`` does.
    """
    n = int(signal.Sample)
    offset = int(file.skip * n * file.dataPrecision * file.dataType)        # :27
    file.fid.seek(offset, 0)
    count = n * file.dataType * n_ms
    if file.dataPrecision == 2:                                             # :28-32
        buf = file.fid.read(count * 2)
        s = np.frombuffer(buf, dtype="<i2").astype(np.float64)
        si, sq = s[0::2], s[1::2]
        return (si - si.mean()) + 1j * (sq - sq.mean())
    buf = file.fid.read(count)                                              # :34
    s = np.frombuffer(buf, dtype=np.int8).astype(np.float64)
    if file.dataType == 2:                                                  # :35-37
        return s[0::2] + 1j * s[1::2]
    return s


def samples_from_bytes(raw_bytes, data_type: int, data_precision: int) -> np.ndarray:
    """Same conversion as :28-38 for a buffer already in memory."""
    if data_precision == 2:
        s = np.frombuffer(raw_bytes, dtype="<i2").astype(np.float64)
        si, sq = s[0::2], s[1::2]
        return (si - si.mean()) + 1j * (sq - sq.mean())
    s = np.frombuffer(raw_bytes, dtype=np.int8).astype(np.float64)
    if data_type == 2:
        return s[0::2] + 1j * s[1::2]
    return s


# --------------------------------------------------------------------------- pieces of the search
def doppler_grid(acq) -> np.ndarray:
    """acquisition.m:42 / :64 -- ``freqMin + freqStep*(freqband-1)``."""
    return acq.freqMin + acq.freqStep * np.arange(int(acq.freqNum), dtype=np.float64)


def carrier_table(signal, acq, coh_ms: int = 1) -> np.ndarray:
    """acquisition.m:41-44 -- ``exp(1i*2*pi*(IF+dopp)*sampleindex./Fs)``.

    Evaluation order kept: ``((1i*2*pi*(IF+dopp)) * n) / Fs``.  For
    ``coh_ms > 1`` (extension) the sample index simply runs on
    over the coherent block, ``n = 1..coh_ms*N``.
    """
    n = np.arange(1, int(signal.Sample) * coh_ms + 1, dtype=np.float64)     # :24
    out = np.empty((int(acq.freqNum), n.size), dtype=np.complex128)
    for b, dopp in enumerate(doppler_grid(acq)):
        out[b, :] = np.exp(((1j * 2 * np.pi * (signal.IF + dopp)) * n) / signal.Fs)   # :43
    return out


def code_replica(signal, prn: int) -> np.ndarray:
    """acquisition.m:49-51 -- nearest-chip (``ceil``) upsampling of the doubled code."""
    n = np.arange(1, int(signal.Sample) + 1, dtype=np.float64)
    ocode = generate_ca_code(prn)
    ocode = np.concatenate((ocode, ocode))                                   # :50
    idx = np.ceil(n * (signal.codeFreqBasis / signal.Fs)).astype(np.int64)   # :51 (1-based)
    return ocode[idx - 1]


def folded_blocks(raw: np.ndarray, carrier: np.ndarray, n: int, k_blocks: int, coh_ms: int) -> np.ndarray:
    """Wipe-off (:56) of every (bin, block), folded to N samples when ``coh_ms > 1``.

    Returns ``temp1`` for each (bin, block): shape (B, K, N).  With
    ``coh_ms == 1`` this is exactly ``raw(block).*carrier(freqband,:)``.
    """
    nb = carrier.shape[0]
    out = np.empty((nb, k_blocks, n), dtype=np.complex128)
    span = n * coh_ms
    for k in range(k_blocks):
        seg = raw[k * span:(k + 1) * span]
        for b in range(nb):
            t1 = seg * carrier[b, :]                                         # :56
            out[b, k, :] = t1 if coh_ms == 1 else t1.reshape(coh_ms, n).sum(axis=0)
    return out


def correlation_surface(raw: np.ndarray, signal, acq, prn: int, *, coh_ms: int = 1,
                        carrier: Optional[np.ndarray] = None,
                        conj_spectra: Optional[np.ndarray] = None,
                        literal: bool = False) -> np.ndarray:
    """acquisition.m:52-61 -- the freqNum x Sample non-coherent power surface of one PRN.

    ``literal=True`` runs the loop body exactly as written (three FFTs per
    (block, bin), ``fft(replica)`` recomputed every time); otherwise the
    PRN-independent ``conj(fft(temp1))`` (``conj_spectra``, shape (B,K,N)) and
    the per-PRN ``fft(replica)`` are computed once -- same operations on the
    same operands, so the values are identical.
    """
    n = int(signal.Sample)
    nb = int(acq.freqNum)
    kk = int(acq.datalen)
    if carrier is None:
        carrier = carrier_table(signal, acq, coh_ms)
    scode = code_replica(signal, prn)                                        # :49-51
    corr = np.zeros((nb, n), dtype=np.float64)                               # :52
    if literal:
        span = n * coh_ms
        for idx in range(kk):                                                # :53
            seg = raw[idx * span:(idx + 1) * span]
            for fb in range(nb):                                             # :54
                replica = scode                                              # :55
                temp1 = seg * carrier[fb, :]                                 # :56
                if coh_ms > 1:
                    temp1 = temp1.reshape(coh_ms, n).sum(axis=0)
                temp2 = np.conj(np.fft.fft(temp1))                           # :57
                temp3 = np.fft.fft(replica)                                  # :58
                corr[fb, :] += np.abs(np.fft.ifft(temp3 * temp2)) ** 2       # :59
        return corr
    if conj_spectra is None:
        conj_spectra = np.conj(np.fft.fft(folded_blocks(raw, carrier, n, kk, coh_ms), axis=-1))
    temp3 = np.fft.fft(scode)
    for idx in range(kk):
        corr += np.abs(np.fft.ifft(temp3[None, :] * conj_spectra[:, idx, :], axis=-1)) ** 2
    return corr


def peak_and_snr(corr: np.ndarray, signal, acq, prn: int, *, snr_threshold_db: float = 12.0,
                 matlab_quirks: bool = True) -> CoarseRow:
    """acquisition.m:62-74 on one PRN's surface."""
    nb, n = corr.shape
    a = np.abs(corr)
    if nb == 1 and matlab_quirks:
        # max(max(.)) of a 1xN row collapses to a scalar first, so both indices come out 1 (A.5).
        fbin, code_phase = 1, 1
        peak = float(a.max())
    else:
        fbin = int(np.argmax(a.max(axis=1))) + 1                             # :62 (first max)
        col_max = a.max(axis=0)
        code_phase = int(np.argmax(col_max)) + 1                             # :63 (first max)
        peak = float(col_max[code_phase - 1])
    doppler = acq.freqMin + acq.freqStep * (fbin - 1)                        # :64
    w = int(math.ceil(signal.Fs / signal.codeFreqBasis))                     # :66
    row = corr[fbin - 1, :]
    lo = row[0:max(code_phase - w, 0)]                                       # 1:codePhase-w
    hi = row[code_phase + w - 1:]                                            # codePhase+w:end
    sel = np.concatenate((lo, hi))
    with np.errstate(divide="ignore", invalid="ignore"):
        noise = float(np.sum(sel ** 2) / sel.size) if sel.size else float("nan")   # :67-68
        snr = float(10.0 * np.log10(np.float64(peak) ** 2 / np.float64(noise)))
    # runner-up: best cell anywhere other than the winning cell (tie-tolerance diagnostic)
    flat = a.copy()
    flat[fbin - 1, code_phase - 1] = -np.inf
    runner = float(flat.max()) if flat.size > 1 else float("nan")
    return CoarseRow(prn=prn, acquired=bool(snr >= snr_threshold_db),        # :70
                     code_phase=code_phase - 1, doppler_bin=fbin - 1, doppler_hz=float(doppler),
                     peak=peak, noise_meansq=noise, snr_db=snr, runner_up=runner)


def coarse_search(raw: np.ndarray, signal, acq, prns: Iterable[int] = range(1, 33), *,
                  coh_ms: int = 1, snr_threshold_db: float = 12.0, literal: bool = False,
                  matlab_quirks: bool = True) -> List[CoarseRow]:
    """acquisition.m:41-80 for the given PRNs (reference: hard-coded 1:32, :47)."""
    n = int(signal.Sample)
    kk = int(acq.datalen)
    carrier = carrier_table(signal, acq, coh_ms)                             # :41-44
    conj_spectra = None
    if not literal:
        conj_spectra = np.conj(np.fft.fft(folded_blocks(raw, carrier, n, kk, coh_ms), axis=-1))
    rows = []
    for prn in prns:                                                         # :47
        corr = correlation_surface(raw, signal, acq, prn, coh_ms=coh_ms, carrier=carrier,
                                   conj_spectra=conj_spectra, literal=literal)
        rows.append(peak_and_snr(corr, signal, acq, prn, snr_threshold_db=snr_threshold_db,
                                 matlab_quirks=matlab_quirks))
    return rows


# --------------------------------------------------------------------------- fine frequency
def fine_frequency(longraw: np.ndarray, file, signal, acq, prn: int, codedelay: int, *, detail: bool = False):
    """acquisition.m:102-121 for one acquired SV.  ``longraw`` = (L+1) ms read as in :89-100.

    ``detail=True`` also returns (1-based peak index, peak magnitude, best other magnitude) so tests can
    apply the floating-point tie tolerance."""
    n = int(signal.Sample)
    ca = generate_ca_code(prn)                                               # :103
    t = np.arange(1, acq.L * n + 1, dtype=np.float64)
    code_idx = np.floor((1.0 / signal.Fs * t) / (1.0 / signal.codeFreqBasis))   # :104
    long_code = ca[(np.fmod(code_idx, signal.codelength)).astype(np.int64)]  # :105 (rem(.)+1, 1-based)
    start = n - int(codedelay)                                               # 1-based (:106)
    seg = longraw[start - 1:start - 1 + acq.L * n] * long_code               # :106
    fftlen = seg.size * int(acq.datalen)                                     # :108
    if file.dataType == 2:
        spec = np.abs(np.fft.fftshift(np.fft.fft(seg, fftlen)))              # :110
    else:
        spec = np.abs(np.fft.fft(seg, fftlen))                               # :112
    half = int(math.ceil(fftlen / 2))                                        # :115
    idx = int(np.argmax(spec[:half * 2])) + 1                                # :116 (1-based, used as is)
    fine = idx * (signal.Fs / fftlen)                                        # :117
    if file.dataType == 2:
        fine = -idx * (signal.Fs / fftlen) + signal.Fs / 2                   # :119
    if detail:
        sel = spec[:half * 2]
        pk = float(sel[idx - 1])
        rest = sel.copy()
        rest[idx - 1] = -np.inf
        return float(fine), idx, pk, float(rest.max())
    return float(fine)


# --------------------------------------------------------------------------- the function itself
def acquisition(file, signal, acq, *, coh_ms: int = 1, snr_threshold_db: float = 12.0,
                literal: bool = False, fine: bool = True, prns: Sequence[int] = tuple(range(1, 33)),
                verbose: bool = False) -> Dict[str, object]:
    """``Acquired = acquisition(file,signal,acq)`` (acquisition.m:1).

    Returns a dict with the five MATLAB fields (1-D float64 arrays, ascending
    PRN, empty when nothing is acquired) plus ``rows``: every PRN's
    :class:`CoarseRow`, acquired or not, for sub-threshold parity checks.
    """
    raw = read_if_block(file, signal, int(acq.datalen) * coh_ms)             # :27-38
    if verbose:
        print("Acquiring... ")                                               # :46
    rows = coarse_search(raw, signal, acq, prns, coh_ms=coh_ms,
                         snr_threshold_db=snr_threshold_db, literal=literal)
    hit = [r for r in rows if r.acquired]                                    # :70-78
    out: Dict[str, object] = {
        "sv": np.array([r.prn for r in hit], dtype=np.float64),
        "SNR": np.array([r.snr_db for r in hit], dtype=np.float64),
        "Doppler": np.array([r.doppler_hz for r in hit], dtype=np.float64),
        "codedelay": np.array([r.code_phase for r in hit], dtype=np.float64),
        "fineFreq": np.array([], dtype=np.float64),
        "rows": rows,
    }
    if verbose:
        for r in hit:                                                        # :76-77
            print(f" SV[{r.prn:2d}] SNR = {r.snr_db:2.2f}, Code phase = {r.code_phase:5d}, "
                  f"Raw Doppler = {int(r.doppler_hz):5d} ")
    if not hit:
        if verbose:
            print("No satellites acquired. Check parameter settings ... ")  # :85
        return out
    if fine:                                                                 # :88-126
        longraw = read_if_block(file, signal, int(acq.L) + 1)
        out["fineFreq"] = np.array(
            [fine_frequency(longraw, file, signal, acq, r.prn, r.code_phase) for r in hit],
            dtype=np.float64)
    return out
