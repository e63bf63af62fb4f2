"""Shared helpers for the parity tests (oracle = checker, libgnssacq = thing under test)."""
from __future__ import annotations

import io
import math
from types import SimpleNamespace

import numpy as np

import oracle
from oracle.synth import SynthSpec, SatSpec, synth_if

# BASELINE.json north_star: indices / decision bit-exact unless the oracle's top two cells are
# closer than the floating-point tolerance; peak and SNR within 1e-4 relative (FP32).
TIE_TOL = 2e-5
METRIC_RTOL = 1e-4


def structs(fs, if_hz, *, data_type=2, data_precision=1, freq_min=-10000.0, freq_step=500.0,
            freq_num=None, datalen=2, skip=0):
    file = SimpleNamespace(fid=None, skip=skip, dataType=data_type, dataPrecision=data_precision)
    signal = oracle.SignalParams(IF=if_hz, Fs=fs)
    acq = oracle.AcqParams(freqStep=freq_step, freqMin=freq_min, freqNum=freq_num, datalen=datalen)
    return file, signal, acq


def small_spec(fs, if_hz, n, *, sats=None, seed=6102, data_type=2, data_precision=1, sigma=16.0):
    if sats is None:
        sats = [SatSpec(3, 990.0, 1683 % n, 1.2, 0.1), SatSpec(7, -3095.0, (n * 2) // 3, 0.9, 1.0),
                SatSpec(22, 1565.0, 17, 1.5, 2.0)]
    return SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=sigma, data_type=data_type,
                     data_precision=data_precision, seed=seed, sats=sats)


def oracle_rows(raw_bytes, file, signal, acq, prns, *, coh_ms=1, matlab_quirks=True):
    file = SimpleNamespace(**vars(file))
    file.fid = io.BytesIO(raw_bytes)
    file.skip = 0
    raw = oracle.read_if_block(file, signal, int(acq.datalen) * coh_ms)
    return oracle.coarse_search(raw, signal, acq, prns, coh_ms=coh_ms, matlab_quirks=matlab_quirks)


def assert_rows_match(gpu_rows, ref_rows, *, thr=12.0, what=""):
    assert len(gpu_rows) == len(ref_rows)
    n_tie = 0
    for g, r in zip(gpu_rows, ref_rows):
        tag = f"{what} PRN {r.prn}"
        assert g.prn == r.prn, tag
        if math.isnan(r.snr_db):
            assert math.isnan(g.snr_db) and not g.acquired, tag
            continue
        tie = r.runner_up >= r.peak * (1.0 - TIE_TOL)
        if tie:
            n_tie += 1
        else:
            assert g.code_phase == r.code_phase, f"{tag}: code phase {g.code_phase} != {r.code_phase}"
            assert g.doppler_bin == r.doppler_bin, f"{tag}: bin {g.doppler_bin} != {r.doppler_bin}"
            assert g.doppler_hz == r.doppler_hz, tag
            assert abs(g.peak - r.peak) <= METRIC_RTOL * r.peak, f"{tag}: peak {g.peak} vs {r.peak}"
            assert abs(g.snr_db - r.snr_db) <= METRIC_RTOL * abs(r.snr_db), f"{tag}: snr {g.snr_db} vs {r.snr_db}"
            if abs(r.snr_db - thr) > METRIC_RTOL * thr:
                assert bool(g.acquired) == r.acquired, tag
    return n_tie
