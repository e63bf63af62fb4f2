"""The configurations BASELINE.json names, at their FULL shapes, against the oracle (`-m gpu`).

* config 1 -- the north_star target: "a full 32-PRN acquisition of the Opensky-shaped block matches
  acquisition.m's code-phase and Doppler indices bit-exactly": 58 MHz / IF 4.58 MHz, 41 bins, K = 20 ms,
  ALL 32 PRNs, every row against the NumPy float64 restatement of acquisition.m:41-80.
* config 3 -- weak signal, 10 ms coherent x 20 non-coherent, 50 Hz step (401 bins, 20 forward bases, shifts to
  +-10 FFT bins), full shape on PRNs present and absent.
* config 5 -- high dynamics, +-50 kHz at 50 Hz (2001 bins, shifts to +-50 FFT bins), 10 ms coherent x 2.

Tolerances: tests/helpers.py (indices / decision exact unless the oracle's top two cells are within 2e-5;
peak and SNR within 1e-4 relative, always).  The oracle works bin block by bin block (helpers.oracle_rows_chunked)
so that the 401 / 2001-bin grids fit in host memory; its arithmetic is coarse_search's.
"""
import numpy as np
import pytest

import oracle
from oracle.synth import SatSpec, SynthSpec, opensky_spec, synth_if
import gnssacq
from gnssacq import api
from helpers import structs, oracle_rows, oracle_rows_chunked, assert_rows_match

pytestmark = pytest.mark.gpu


def test_config1_opensky_all_32_prns_full_depth():
    """BASELINE config 1 / north_star target, nothing reduced: 32 PRNs x 41 bins x 20 blocks of 58 000 samples."""
    file, signal, acq = gnssacq.initParameters(shape="opensky")
    assert (int(signal.Sample), int(acq.freqNum), int(acq.datalen)) == (58000, 41, 20)
    spec = opensky_spec()
    raw_b = synth_if(spec, 0, 20)
    prns = list(range(1, 33))
    ref = oracle_rows(raw_b, file, signal, acq, prns)
    with api.Searcher(gnssacq.config_from_structs(file, signal, acq, prns=prns)) as s:
        rows = s.search(raw_b)
        again = s.search(raw_b)
    ties = assert_rows_match(rows, ref, what="config 1, 32 PRNs, K=20")
    assert ties <= 1
    assert [bytes(r) for r in rows] == [bytes(r) for r in again]            # idempotent
    # the synthetic truth table (from the reference's own Acquired_Opensky_5000.mat) is recovered exactly
    got = {r.prn: r for r in rows if r.acquired}
    for sat in spec.sats:
        assert sat.prn in got and got[sat.prn].code_phase == sat.codedelay
        assert abs(got[sat.prn].doppler_hz - sat.doppler_hz) <= 250.0
    assert {r.prn for r in ref if r.acquired} == set(got)


def _weak_spec(dopplers):
    """Opensky front end, four weak satellites (0.06-0.12 LSB against sigma = 16 LSB per component)."""
    sats = [SatSpec(3, dopplers[0], 3683, 0.10, 0.3), SatSpec(16, dopplers[1], 26051, 0.12, 1.1),
            SatSpec(22, dopplers[2], 2610, 0.08, 2.0), SatSpec(31, dopplers[3], 39064, 0.08, 0.7)]
    return SynthSpec(sats=sats, seed=6102 + 3)


def test_config3_weak_signal_full_shape():
    """BASELINE config 3: 10 ms coherent x 20 non-coherent, -10 kHz : 50 Hz : +10 kHz (401 bins), N = 58 000."""
    file, signal, acq = structs(58e6, 4.58e6, datalen=20, freq_min=-10000.0, freq_step=50.0, freq_num=401)
    spec = _weak_spec([990.0, -305.0, 1565.0, 9870.0])
    raw_b = synth_if(spec, 0, 200)
    prns = [3, 8, 31]                                                         # present, absent, present near the grid edge
    ref = oracle_rows_chunked(raw_b, file, signal, acq, prns, coh_ms=10, chunk_bins=20)
    cfg = gnssacq.config_from_structs(file, signal, acq, prns=prns, coh_ms=10)
    with api.Searcher(cfg) as s:
        rows = s.search(raw_b)
        assert s.last_stats.n_bases == 20
    assert_rows_match(rows, ref, what="config 3")
    by = {r.prn: r for r in rows}
    assert by[3].acquired and by[3].code_phase == 3683 and abs(by[3].doppler_hz - 990.0) <= 25.0
    assert by[31].acquired and by[31].code_phase == 39064 and abs(by[31].doppler_hz - 9870.0) <= 25.0


def test_config5_high_dynamics_full_shape():
    """BASELINE config 5: +-50 kHz at 50 Hz (2001 bins), 10 ms coherent x 2, N = 58 000."""
    file, signal, acq = structs(58e6, 4.58e6, datalen=2, freq_min=-50000.0, freq_step=50.0, freq_num=2001)
    spec = _weak_spec([-47310.0, 12345.0, 48020.0, -60.0])
    spec.sats = [SatSpec(s.prn, s.doppler_hz, s.codedelay, 4.0 * s.amplitude, s.phase) for s in spec.sats]
    raw_b = synth_if(spec, 0, 20)
    prns = [3, 22]                                                            # near -50 kHz and near +50 kHz
    ref = oracle_rows_chunked(raw_b, file, signal, acq, prns, coh_ms=10, chunk_bins=100)
    cfg = gnssacq.config_from_structs(file, signal, acq, prns=prns, coh_ms=10)
    with api.Searcher(cfg) as s:
        rows = s.search(raw_b)
        assert s.last_stats.n_bases == 20
    assert_rows_match(rows, ref, what="config 5")
    by = {r.prn: r for r in rows}
    assert by[3].code_phase == 3683 and abs(by[3].doppler_hz + 47310.0) <= 25.0
    assert by[22].code_phase == 2610 and abs(by[22].doppler_hz - 48020.0) <= 25.0
