"""Tracking correlators through the C ABI against oracle/tracking_ref.py.  `-m gpu`.

Bar: float64 sums; |dI|, |dQ| <= 1e-9 of the channel's largest |I|,|Q| (summation order and sincos ulps), i.e.
every code-phase sample falls on the same chip as in the oracle."""
import numpy as np
import pytest

import gnssacq
from gnssacq import api
from oracle import tracking_ref as tr
from oracle.synth import SatSpec, SynthSpec, synth_if
from helpers import structs

pytestmark = pytest.mark.gpu

EPL = [-0.5, 0.0, 0.5]                                  # trackingCT.m:24 with track.CorrelatorSpacing = 0.5
BANK25 = [round(0.6 - 0.05 * i, 2) for i in range(25)]  # trackingCT_POS_updated_multicorrelator.m:41


def _cfg(fs, if_hz, data_type, precision):
    file, signal, acq = structs(fs, if_hz, data_type=data_type, data_precision=precision, datalen=2)
    return gnssacq.config_from_structs(file, signal, acq, prns=[1])


def _check(gi, gq, oi, oq, tag):
    scale = max(np.abs(oi).max(), np.abs(oq).max(), 1.0)
    assert np.abs(gi - oi).max() <= 1e-9 * scale, f"{tag}: I {np.abs(gi - oi).max()} of {scale}"
    assert np.abs(gq - oq).max() <= 1e-9 * scale, f"{tag}: Q {np.abs(gq - oq).max()} of {scale}"


@pytest.mark.parametrize("fs,if_hz,data_type,precision,taps", [
    (58e6, 4.58e6, 2, 1, EPL), (26e6, 0.0, 2, 1, BANK25), (58e6, 4.58e6, 1, 1, EPL), (6e6, 1.25e6, 2, 2, BANK25),
    (58e6, 4.58e6, 2, 1, BANK25),
])
def test_correlators_against_oracle(fs, if_hz, data_type, precision, taps):
    n = int(fs * 1e-3)
    rng = np.random.default_rng(int(fs / 1e6) + 10 * data_type + precision)
    sats = [SatSpec(3, 990.0, 1683 % n, 3.0, 0.1), SatSpec(22, -2310.0, (2 * n) // 3, 2.0, 1.0)]
    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=8.0, data_type=data_type, data_precision=precision,
                     seed=31, sats=sats)
    raw = synth_if(spec, 0, 4)
    x_all = tr.samples_of(raw, data_type, 1) if precision == 1 else None
    chans = []
    for prn, f_d, cd in [(3, 990.0, 1683 % n), (22, -2310.0, (2 * n) // 3), (9, 120.0, 5)]:
        for _ in range(2):                                             # random loop states around the truth
            code_hz = 1.023e6 + rng.uniform(-3, 3)
            rem_chip = rng.uniform(-0.2, 0.9)
            ns = tr.num_samples(code_hz, fs, rem_chip)
            chans.append(api.Channel(prn=prn, num_samples=ns, sample_offset=int(n - cd + 1 + rng.integers(0, n)),
                                     carrier_hz=if_hz + f_d + rng.uniform(-20, 20), rem_phase=rng.uniform(-6, 6),
                                     code_hz=code_hz, rem_chip=rem_chip))
    with api.Searcher(_cfg(fs, if_hz, data_type, precision)) as s:
        s.track_load(raw)
        gi, gq = s.correlate(chans, taps)
        gi2, gq2 = s.correlate(chans, taps)
        assert np.array_equal(gi, gi2) and np.array_equal(gq, gq2)      # fixed-order reduction: deterministic
        one_i, one_q = s.correlate(chans[:1], taps[:1])                 # batching changes nothing
        assert one_i[0, 0] == gi[0, 0] and one_q[0, 0] == gq[0, 0]
    bps = data_type * precision
    for c, ch in enumerate(chans):
        seg = raw[ch.sample_offset * bps:(ch.sample_offset + ch.num_samples) * bps]
        x = tr.samples_of(seg, data_type, precision)                    # int16: per-integration DC removal (:90-92)
        oi, oq = tr.correlate(x, fs, ch.prn, ch.carrier_hz, ch.rem_phase, ch.code_hz, ch.rem_chip, taps)
        _check(gi[c], gq[c], oi, oq, f"fs={fs} type={data_type}/{precision} channel {c}")


def test_closed_loop_tracking_matches_the_oracle_loop():
    """trackingCT.m:70-150 for two channels over 40 ms: the loop filters run on the host (restated in the oracle
    module), the correlations come from the GPU; the oracle runs the same loop on its own correlations.  The two
    trajectories must stay together (same numSample every period) and lock."""
    fs, if_hz, n = 26e6, 0.0, 26000
    truth = [(5, 1500.0, 4000, 5.0), (17, -2750.0, 15000, 4.0)]
    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=6.0, data_type=2, data_precision=1, seed=5,
                     sats=[SatSpec(p, d, cd, a, 0.4) for p, d, cd, a in truth])
    raw = synth_if(spec, 0, 43)
    x_all = tr.samples_of(raw, 2, 1)
    gpu = [tr.ChannelState(prn=p, carrier_basis_hz=if_hz + d + 8.0, carrier_hz=if_hz + d + 8.0, sample_pos=n - cd + 1)
           for p, d, cd, a in truth]
    cpu = [tr.ChannelState(prn=p, carrier_basis_hz=if_hz + d + 8.0, carrier_hz=if_hz + d + 8.0, sample_pos=n - cd + 1)
           for p, d, cd, a in truth]
    with api.Searcher(_cfg(fs, if_hz, 2, 1)) as s:
        s.track_load(raw)
        for ms in range(40):
            ns = [tr.num_samples(st.code_hz, fs, st.rem_chip) for st in gpu]
            chans = [api.Channel(prn=st.prn, num_samples=k, sample_offset=st.sample_pos, carrier_hz=st.carrier_hz,
                                 rem_phase=st.rem_phase, code_hz=st.code_hz, rem_chip=st.rem_chip)
                     for st, k in zip(gpu, ns)]
            gi, gq = s.correlate(chans, EPL)
            for c, (sg, sc) in enumerate(zip(gpu, cpu)):
                kc = tr.num_samples(sc.code_hz, fs, sc.rem_chip)
                assert kc == ns[c], f"ms {ms} channel {c}: numSample diverged"
                oi, oq = tr.correlate(x_all[sc.sample_pos:sc.sample_pos + kc], fs, sc.prn, sc.carrier_hz, sc.rem_phase,
                                      sc.code_hz, sc.rem_chip, EPL)
                _check(gi[c], gq[c], oi, oq, f"ms {ms} channel {c}")
                tr.close_loops(sg, gi[c], gq[c], ns[c], fs)
                tr.close_loops(sc, oi, oq, kc, fs)
    for st, (p, d, cd, a) in zip(gpu, truth):
        pw = np.array([np.hypot(r["P_i"], r["P_q"]) for r in st.history[15:]])
        assert np.median(pw) > 0.8 * a * n
        assert abs(np.mean([r["carrier_hz"] for r in st.history[15:]]) - (if_hz + d)) < 15.0


def test_correlate_argument_errors():
    with api.Searcher(_cfg(6e6, 1.25e6, 2, 1)) as s:
        ch = api.Channel(prn=1, num_samples=6000, sample_offset=0, carrier_hz=1.25e6, rem_phase=0.0, code_hz=1.023e6, rem_chip=0.0)
        with pytest.raises(gnssacq.GnssAcqError) as e:
            s.correlate([ch], EPL)                                       # nothing loaded
        assert e.value.code == -7
        s.track_load(bytes(2 * 6000 * 2))
        s.correlate([ch], EPL)
        for bad in (dict(prn=0), dict(prn=38), dict(num_samples=0), dict(sample_offset=-1), dict(sample_offset=6001),
                    dict(code_hz=0.0)):
            b = api.Channel(**{**{f: getattr(ch, f) for f, _ in api.Channel._fields_}, **bad})
            with pytest.raises(gnssacq.GnssAcqError):
                s.correlate([b], EPL)
        with pytest.raises(gnssacq.GnssAcqError):
            s.correlate([ch], [0.0] * 33)


@pytest.mark.parametrize("fs,if_hz,data_type,precision,periods", [(26e6, 0.0, 2, 1, 60), (58e6, 4.58e6, 2, 1, 25),
                                                                   (6e6, 1.25e6, 2, 2, 40), (6e6, 1.25e6, 1, 1, 40)])
def test_device_side_tracking_loop_against_the_oracle_loop(fs, if_hz, data_type, precision, periods):
    """gnssacq_track: trackingCT.m:70-172 closed on the GPU (one cluster per channel, no host round trip per ms).
    Period by period against the oracle's loop: numSample and the file position identical, sums to 1e-7 of the
    prompt magnitude, NCO frequencies to 1e-6 Hz; and the loops lock."""
    n = int(fs * 1e-3)
    truth = [(5, 1500.0, (n * 2) // 13, 5.0), (17, -2750.0, (n * 7) // 12, 4.0), (30, 420.0, 11, 6.0)]
    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=6.0, data_type=data_type, data_precision=precision,
                     seed=5, sats=[SatSpec(p, d, cd, a, 0.4) for p, d, cd, a in truth])
    raw = synth_if(spec, 0, periods + 3)
    start = [api.Channel(prn=p, num_samples=0, sample_offset=n - cd + 1, carrier_hz=if_hz + d + 8.0, rem_phase=0.0,
                         code_hz=1.023e6, rem_chip=0.0) for p, d, cd, a in truth]
    with api.Searcher(_cfg(fs, if_hz, data_type, precision)) as s:
        s.track_load(raw)
        recs = s.track(start, periods)
        again = s.track(start, periods)
        assert [bytes(r) for ch in recs for r in ch] == [bytes(r) for ch in again for r in ch]       # deterministic
        with pytest.raises(gnssacq.GnssAcqError) as e:                                                # :107-111
            s.track(start, periods + 10)
        assert e.value.code == -3
    bps = data_type * precision
    for c, (p, d, cd, a) in enumerate(truth):
        st = tr.ChannelState(prn=p, carrier_basis_hz=if_hz + d + 8.0, carrier_hz=if_hz + d + 8.0, sample_pos=n - cd + 1)
        for i, r in enumerate(recs[c]):
            ns = tr.num_samples(st.code_hz, fs, st.rem_chip)
            assert r.num_samples == ns, f"channel {c} period {i}: numSample {r.num_samples} != {ns}"
            x = tr.samples_of(raw[st.sample_pos * bps:(st.sample_pos + ns) * bps], data_type, precision)
            oi, oq = tr.correlate(x, fs, p, st.carrier_hz, st.rem_phase, st.code_hz, st.rem_chip, EPL)
            tr.close_loops(st, oi, oq, ns, fs)
            scale = max(np.hypot(oi[1], oq[1]), 1.0)
            got = np.array([r.E_i, r.P_i, r.L_i, r.E_q, r.P_q, r.L_q])
            assert np.abs(got - np.concatenate([oi, oq])).max() <= 1e-7 * scale, f"channel {c} period {i}"
            assert r.sample_end == st.sample_pos
            assert abs(r.code_hz - st.code_hz) <= 1e-6 and abs(r.carrier_hz - st.carrier_hz) <= 1e-6
            assert abs(r.rem_chip - st.rem_chip) <= 1e-9 and abs(r.rem_phase - st.rem_phase) <= 1e-7
        tail = recs[c][15:]
        if data_type == 2:                                   # (a real-sampled signal carries half the power per sideband)
            assert np.median([np.hypot(r.P_i, r.P_q) for r in tail]) > 0.8 * a * n
        assert abs(np.mean([r.carrier_hz for r in tail]) - (if_hz + d)) < 15.0


def test_trackingct_twin_on_a_synthetic_recording():
    """gnssacq.trackingCT (stage 1 of trackingCT.m) on a synthetic 6 MHz recording: locks on every acquired SV, the
    per-period fields have the reference's meaning, C/N0 matches the generator's signal and noise levels, and the
    bit-edge index points at the generator's 20 ms data-bit grid."""
    from oracle.synth import VirtualFile
    fs, if_hz, n = 6e6, 1.25e6, 6000
    truth = [(3, 990.0, 4800, 5.0), (22, -2310.0, 5100, 6.0)]          # code delays near N: bit edges late in a period
    sigma = 5.0
    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=sigma, data_type=2, data_precision=1, seed=77,
                     sats=[SatSpec(p, d, cd, a, 0.2) for p, d, cd, a in truth])
    file, signal, acq = structs(fs, if_hz, datalen=2, skip=3)
    file.fid = VirtualFile(spec)
    signal.ms, signal.Sample = 1e-3, n
    track = gnssacq.trackParameters()
    track.msToProcessCT_1ms = 700
    acquired = {"sv": np.array([p for p, *_ in truth], float), "codedelay": np.array([cd for _, _, cd, _ in truth], float),
                "fineFreq": np.array([if_hz + d + 6.0 for _, d, _, _ in truth])}
    res, cn0, countinx = gnssacq.trackingCT(file, signal, track, acquired)
    gnssacq.release_all()
    assert sorted(res) == [3, 22] and cn0.shape == (35, 2) and countinx.shape == (2,)
    for c, (p, d, cd, a) in enumerate(truth):
        r = res[p]
        assert all(len(r[k]) == 700 for k in r)
        pw = np.hypot(r["P_i"][50:], r["P_q"][50:])
        assert np.median(pw) > 0.85 * a * n
        assert abs(np.mean(r["carrierFreq"][100:]) - (if_hz + d)) < 5.0
        assert abs(np.mean(r["codeFreq"][100:]) - 1.023e6) < 2.0          # (the generator has no code Doppler)
        assert np.all(np.abs(r["delayValue"]) <= 1) and np.all(r["numSample"] == n + r["delayValue"])
        # per-satellite cumulative sum: a documented, deliberate deviation from trackingCT.m:161, whose linear index
        # into the n_sv x n_ms delayValue matrix mixes the satellites (gnssacq/tracking.py)
        assert np.allclose(r["codedelay"], cd + np.cumsum(r["delayValue"]))
        bps = 2
        assert r["absoluteSample"][0] == ((3 * n) + (n - cd + 1) + r["numSample"][0]) * bps       # ftell after the first read
        # C/N0 (trackingCT.m:121-133): the formula itself is pinned on the CPU (tests/test_tracking_host.py); here
        # the wiring -- 35 blocks of 20 periods from this channel's prompt sums.  (At 6 samples per chip the +-1
        # sample code-phase steps modulate |P| by a few per cent, which caps what the moment method reports far
        # below the generator's 65 dB-Hz.)
        assert np.array_equal(cn0[:, c], gnssacq.cn0_estimates(r["P_i"], r["P_q"]))
        assert 40.0 < np.nanmedian(cn0[:, c]) < 70.0
        # the generator flips data bits on file-ms multiples of 20: (cd - 1) samples into tracking period 20m - 3
        # (skip = 3 ms), so the first fully flipped period is 20m - 2 (1-based) and countinx = mod(i, 20) - 1 = 17,
        # unless no bit flipped between periods 600 and 682 (then 0)
        assert countinx[c] in (17.0, 0.0)
    assert (countinx == 17.0).any()
