"""``Acquired = acquisition(file, signal, acq)`` -- Python twin of the MATLAB wrapper.

Mirrors the call signature and result fields of
``SDR_MATLAB-main/acqtckpos/acquisition.m:1``: the three structs go in
(attribute access: ``file.fid``, ``signal.Sample`` ...), a dict with ``sv``,
``SNR``, ``Doppler``, ``codedelay``, ``fineFreq`` (1-D float64, ascending PRN,
empty when nothing is acquired) comes out, and the same progress lines are
printed.  File I/O stays on the host side of the boundary exactly as in the
MATLAB wrapper (``fseek``/``fread`` of raw bytes, acquisition.m:27-38); every
numeric step of the coarse search runs in ``libgnssacq.so`` on the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import api

_searchers: Dict[tuple, api.Searcher] = {}


def config_from_structs(file, signal, acq, *, coh_ms: int = 1, prns: Sequence[int] = tuple(range(1, 33)),
                        snr_threshold_db: float = 12.0, device: int = -1, cluster_ctas: int = 0,
                        threads: int = 0, keep_surface: bool = False, exchange: int = 0,
                        work_split: int = 0) -> api.Config:
    """Map exactly the fields acquisition.m reads onto ``gnssacq_config``."""
    return api.make_config(
        fs_hz=float(signal.Fs), if_hz=float(signal.IF), code_hz=float(signal.codeFreqBasis),
        samples_per_ms=int(signal.Sample), data_type=int(file.dataType),
        data_precision=int(file.dataPrecision), freq_min_hz=float(acq.freqMin),
        freq_step_hz=float(acq.freqStep), freq_num=int(acq.freqNum), noncoh_blocks=int(acq.datalen),
        coh_ms=coh_ms, prns=prns, snr_threshold_db=snr_threshold_db, device=device,
        cluster_ctas=cluster_ctas, threads=threads, keep_surface=keep_surface, exchange=exchange, work_split=work_split)


def _key(cfg: api.Config) -> tuple:
    return bytes(cfg)


def get_searcher(cfg: api.Config) -> api.Searcher:
    """One persistent handle per distinct configuration (the MEX gateway keeps one in a static)."""
    k = _key(cfg)
    s = _searchers.get(k)
    if s is None:
        s = api.Searcher(cfg)
        _searchers[k] = s
    return s


def release_all() -> None:
    for s in _searchers.values():
        s.close()
    _searchers.clear()


def read_if_bytes(file, signal, n_ms: int) -> bytes:
    """acquisition.m:27,29/34 -- absolute seek, then read ``n_ms`` ms of raw samples."""
    nbytes = int(signal.Sample) * int(file.dataType) * int(file.dataPrecision)
    file.fid.seek(int(file.skip) * nbytes, 0)
    return file.fid.read(nbytes * n_ms)


def acquisition(file, signal, acq, *, coh_ms: int = 1, verbose: bool = True, fine: bool = True,
                return_rows: bool = False, searcher: Optional[api.Searcher] = None):
    cfg = searcher.cfg if searcher is not None else config_from_structs(file, signal, acq, coh_ms=coh_ms)
    s = searcher if searcher is not None else get_searcher(cfg)
    raw = read_if_bytes(file, signal, int(acq.datalen) * coh_ms)
    if verbose:
        print("Acquiring... ")                                                   # acquisition.m:46
    rows = s.search(raw)
    hit = [r for r in rows if r.acquired]                                        # :70-74
    Acquired = {
        "sv": np.array([r.prn for r in hit], dtype=np.float64),
        "SNR": np.array([r.snr_db for r in hit], dtype=np.float64),
        "Doppler": np.array([r.doppler_hz for r in hit], dtype=np.float64),
        "codedelay": np.array([r.code_phase for r in hit], dtype=np.float64),
        "fineFreq": np.array([], dtype=np.float64) if fine else np.full(len(hit), np.nan),
    }
    if verbose:
        for r in hit:                                                            # :76-77
            print(f" SV[{r.prn:2d}] SNR = {r.snr_db:2.2f}, Code phase = {r.code_phase:5d}, "
                  f"Raw Doppler = {int(r.doppler_hz):5d} ")
        if not hit:
            print("No satellites acquired. Check parameter settings ... ")      # :85
    if fine and hit:                                                             # :88-126, on the GPU
        if verbose:
            print("Now refining Doppler freq... ")
        long_raw = read_if_bytes(file, signal, int(acq.L) + 1)                   # :89-100
        Acquired["fineFreq"] = s.fine_frequency(long_raw, int(acq.L), [r.prn for r in hit],
                                                [r.code_phase for r in hit])
        if verbose:
            for r, ff in zip(hit, Acquired["fineFreq"]):                         # :123-125
                print(f" SV[{r.prn:2d}] SNR = {r.snr_db:2.2f}, Code phase = {r.code_phase:5d}, "
                      f"Raw Doppler = {int(r.doppler_hz):5d}, Fine Doppler = {ff - signal.IF:5f} ")
    if return_rows:
        return Acquired, rows
    return Acquired
