"""Pins the oracle (CPU, runs everywhere).  PARITY UNPINNED against MATLAB itself (see oracle/__init__.py);
these are the independent checks SURVEY.md 8(c) lists."""
import io
import json
import os

import numpy as np
import pytest

import oracle
from oracle.acquisition_ref import (carrier_table, code_replica, correlation_surface, peak_and_snr,
                                    folded_blocks)
from oracle.cacode import IS_GPS_200_FIRST10_OCTAL
from oracle.synth import SatSpec, VirtualFile, synth_if, synth_samples
from helpers import structs, small_spec, oracle_rows

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_ca_code_is_gps_200_octal():
    for prn in range(1, 33):
        c = oracle.generate_ca_code(prn)
        assert c.shape == (1023,) and set(np.unique(c)) == {-1.0, 1.0}
        bits = "".join("1" if v > 0 else "0" for v in c[:10])      # chip +1 <-> logic 1
        assert int(bits, 2) == IS_GPS_200_FIRST10_OCTAL[prn - 1], prn


def test_ca_code_autocorrelation_values():
    for prn in (1, 13, 32):
        c = oracle.generate_ca_code(prn)
        ac = np.rint(np.fft.ifft(np.abs(np.fft.fft(c)) ** 2).real).astype(int)
        assert ac[0] == 1023 and set(np.unique(ac[1:])) <= {63, -1, -65}


def test_code_replica_matches_exact_rational_ceil():
    from fractions import Fraction
    for fs in (58e6, 26e6, 6e6):
        _, s, _ = structs(fs, 0.0)
        n = int(s.Sample)
        ca = oracle.generate_ca_code(5)
        got = code_replica(s, 5)
        ratio = Fraction(1023000, int(fs))
        exact = np.array([ca[(-((-k * ratio.numerator) // ratio.denominator) - 1) % 1023] for k in range(1, n + 1)])
        # double arithmetic (what MATLAB does, acquisition.m:51) can only differ from the exact rational
        # ceil where n*fc/Fs is an integer; for the two real front ends it never does.
        diff = np.nonzero(got != exact)[0] + 1
        assert all((k * ratio).denominator == 1 for k in diff), fs
        if fs in (58e6, 26e6):
            assert diff.size == 0, fs


def test_surface_against_bruteforce_time_domain():
    """acquisition.m:56-59 computes |sum_n conj(x_w[n]) c[(n+m) mod N]|^2: check random cells."""
    file, signal, acq = structs(6e6, 1.25e6, datalen=2, freq_min=-1500.0, freq_step=500.0)
    n = int(signal.Sample)
    raw_b = synth_if(small_spec(6e6, 1.25e6, n), 0, 2)
    file.fid = io.BytesIO(raw_b)
    raw = oracle.read_if_block(file, signal, 2)
    corr = correlation_surface(raw, signal, acq, 3)
    lit = correlation_surface(raw, signal, acq, 3, literal=True)
    assert np.array_equal(corr, lit)                      # cached form == literal loop body
    car = carrier_table(signal, acq).astype(np.clongdouble)
    c = code_replica(signal, 3).astype(np.longdouble)
    rng = np.random.default_rng(1)
    for _ in range(40):
        b, m = int(rng.integers(0, acq.freqNum)), int(rng.integers(0, n))
        tot = np.longdouble(0)
        for k in range(2):
            xw = raw[k * n:(k + 1) * n].astype(np.clongdouble) * car[b]
            tot += np.abs(np.sum(np.conj(xw) * np.roll(c, -m))) ** 2
        assert abs(float(tot) - corr[b, m]) <= 1e-9 * corr.max()


def test_truth_recovery_and_schema():
    file, signal, acq = structs(6e6, 1.25e6, datalen=4)
    n = int(signal.Sample)
    spec = small_spec(6e6, 1.25e6, n)
    file.fid = VirtualFile(spec)
    file.skip = 3
    out = oracle.acquisition(file, signal, acq, fine=False)
    got = {int(p): (int(cd), float(d)) for p, cd, d in zip(out["sv"], out["codedelay"], out["Doppler"])}
    for s in spec.sats:
        assert s.prn in got
        assert got[s.prn][0] == s.codedelay
        assert abs(got[s.prn][1] - s.doppler_hz) <= 250.0
    for k in ("sv", "SNR", "Doppler", "codedelay", "fineFreq"):
        assert out[k].dtype == np.float64 and out[k].ndim == 1
    assert list(out["sv"]) == sorted(out["sv"])


def test_reference_saved_results_schema():
    """tests/golden/acquired_*.json are the reference's own saved outputs (make_golden.py)."""
    for name, n in (("acquired_opensky_5000.json", 58000), ("nacquired_urban_5000.json", 26000)):
        g = json.load(open(os.path.join(GOLDEN, name)))
        assert set(g["fields"]) == {"sv", "SNR", "Doppler", "codedelay", "fineFreq"}
        k = len(g["sv"])
        assert all(len(g[f]) == k for f in g["fields"])
        assert all(0 <= c < n and float(c).is_integer() for c in g["codedelay"])
        assert all(d % 500 == 0 and abs(d) <= 10000 for d in g["Doppler"])
        assert all(abs(ff - g["IF"] - d) <= 500 for ff, d in zip(g["fineFreq"], g["Doppler"]))
        assert all(s >= 12.0 for s in g["SNR"])


def test_oracle_regression_rows():
    """Oracle output on a committed seeded input must not drift (tests/golden/small_rows.json)."""
    g = json.load(open(os.path.join(GOLDEN, "small_rows.json")))
    file, signal, acq = structs(g["fs"], g["if"], datalen=g["datalen"])
    raw_b = synth_if(small_spec(g["fs"], g["if"], int(signal.Sample), seed=g["seed"]), 0, g["datalen"])
    rows = oracle_rows(raw_b, file, signal, acq, g["prns"])
    for r, e in zip(rows, g["rows"]):
        assert (r.prn, r.code_phase, r.doppler_bin, r.acquired) == (e["prn"], e["code_phase"], e["doppler_bin"], e["acquired"])
        assert abs(r.peak - e["peak"]) <= 1e-9 * e["peak"] and abs(r.snr_db - e["snr_db"]) <= 1e-9 * abs(e["snr_db"])


def test_coherent_fold_equals_long_correlation():
    """folding M wiped-off ms before the FFT == correlating M*N samples against the wrapped code."""
    file, signal, acq = structs(6e6, 1.25e6, datalen=1, freq_min=0.0, freq_step=500.0, freq_num=2)
    n, m_coh = int(signal.Sample), 3
    raw_b = synth_if(small_spec(6e6, 1.25e6, n), 0, m_coh)
    file.fid = io.BytesIO(raw_b)
    raw = oracle.read_if_block(file, signal, m_coh)
    corr = correlation_surface(raw, signal, acq, 3, coh_ms=m_coh)
    car = carrier_table(signal, acq, m_coh)
    c = np.tile(code_replica(signal, 3), m_coh)
    for b in range(2):
        xw = raw * car[b]
        long_corr = np.abs(np.fft.ifft(np.fft.fft(c) * np.conj(np.fft.fft(xw)))) ** 2
        assert np.allclose(long_corr[:n], corr[b], rtol=1e-9, atol=1e-9 * corr.max())


def test_peak_search_tie_and_single_bin_quirk():
    _, signal, acq = structs(6e6, 0.0, freq_num=3)
    n = int(signal.Sample)
    corr = np.ones((3, n))
    corr[2, 100] = corr[1, 40] = 5.0                        # exact tie in two bins
    r = peak_and_snr(corr, signal, acq, 1)
    assert (r.doppler_bin, r.code_phase) == (1, 40)         # first bin; first column over all bins
    acq1 = oracle.AcqParams(freqNum=1)
    r1 = peak_and_snr(corr[2:3], signal, acq1, 1)
    assert (r1.doppler_bin, r1.code_phase) == (0, 0)        # max(max(.)) of a 1xN row
    r2 = peak_and_snr(corr[2:3], signal, acq1, 1, matlab_quirks=False)
    assert r2.code_phase == 100


def test_snr_window_is_clipped_not_wrapped():
    _, signal, acq = structs(6e6, 0.0, freq_num=1)
    n, w = int(signal.Sample), 6
    corr = np.full((2, n), 2.0)
    corr[0, 2] = 50.0
    acq2 = oracle.AcqParams(freqNum=2)
    r = peak_and_snr(corr, signal, acq2, 1)
    count = n - (2 + 1) - w + 1                             # only codePhase+w..N survives
    assert abs(r.noise_meansq - 4.0) < 1e-12
    assert r.code_phase == 2 and count > 0
