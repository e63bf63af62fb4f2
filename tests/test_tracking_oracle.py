"""CPU pins of the tracking-correlator oracle (oracle/tracking_ref.py, trackingCT.m:75-150)."""
import numpy as np

from oracle import tracking_ref as tr
from oracle.cacode import generate_ca_code
from oracle.synth import SatSpec, SynthSpec, synth_if


def test_loop_coefficients_and_sample_count():
    tau1, tau2 = tr.calc_loop_coef(2.0, 0.707, 0.1)                     # calcLoopCoef.m:41-45, track.DLL* defaults
    wn = 2.0 * 8 * 0.707 / (4 * 0.707 ** 2 + 1)
    assert tau1 == 0.1 / wn ** 2 and tau2 == 2 * 0.707 / wn
    assert tr.num_samples(1.023e6, 58e6, 0.0) == 58000                  # trackingCT.m:78 at nominal rates
    assert tr.num_samples(1.023e6, 26e6, 0.0) == 26000
    assert tr.num_samples(1.023e6 + 2.0, 58e6, 0.3) in (57982, 57983)


def test_correlate_is_the_brute_force_sum():
    """The vectorised restatement against a per-sample loop written straight from trackingCT.m:96-117."""
    rng = np.random.default_rng(7)
    fs, n, prn = 6e6, 601, 9
    x = rng.integers(-100, 100, n) + 1j * rng.integers(-100, 100, n)
    f, ph, fc, rc, spacing = 1.25e6 + 431.0, 0.7, 1.023e6 - 1.5, 0.37, [-0.5, 0.0, 0.5]
    gi, gq = tr.correlate(x, fs, prn, f, ph, fc, rc, spacing)
    ca = generate_ca_code(prn)
    code = [ca[-1]] + list(ca) + [ca[0]]
    for t, sp in enumerate(spacing):
        si = sq = 0.0
        for k in range(n):
            tt = (0.0 + sp + rc) + (fc / fs) * k
            chip = code[int(np.ceil(tt))]                                # 1-based ceil(t)+1 -> 0-based ceil(t)
            m = x[k] * np.exp(1j * ((2 * np.pi * (f * (k / fs))) + ph))
            si += chip * m.imag
            sq += chip * m.real
        assert abs(gi[t] - si) <= 1e-9 * max(1.0, abs(si)) and abs(gq[t] - sq) <= 1e-9 * max(1.0, abs(sq))


def test_closed_loop_locks_on_a_synthetic_satellite():
    """trackingCT.m:24-150 on the restatement: starting from the acquisition result the prompt power stays high,
    the early/late powers balance and the carrier settles near the true Doppler."""
    fs, if_hz, n, prn, doppler, codedelay = 6e6, 1.25e6, 6000, 7, 1410.0, 2345
    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=n, sigma=4.0, data_type=2, data_precision=1, seed=11,
                     sats=[SatSpec(prn, doppler, codedelay, 6.0, 0.3)])
    raw = synth_if(spec, 0, 62)
    x_all = tr.samples_of(raw, 2, 1)
    st = tr.ChannelState(prn=prn, carrier_basis_hz=if_hz + doppler + 12.0, carrier_hz=if_hz + doppler + 12.0,
                         sample_pos=n - codedelay + 1)                    # trackingCT.m:60 (Sample - AcqCodeDelay + 1)
    # the synthetic signal sits at -(IF + fd): wipe-off with +f like the reference
    for _ in range(60):
        ns = tr.num_samples(st.code_hz, fs, st.rem_chip)
        x = x_all[st.sample_pos:st.sample_pos + ns]
        i, q = tr.correlate(x, fs, prn, st.carrier_hz, st.rem_phase, st.code_hz, st.rem_chip, [-0.5, 0.0, 0.5])
        tr.close_loops(st, i, q, ns, fs)
    h = st.history[20:]
    p = np.array([np.hypot(r["P_i"], r["P_q"]) for r in h])
    # coherent gain ~ amplitude * n; the generator flips nav bits on file-ms edges, i.e. inside an integration that
    # starts at the code edge, so a couple of periods lose power -- not a loss of lock
    assert np.median(p) > 0.85 * 6.0 * n and (p < 0.5 * np.median(p)).sum() <= 3
    assert abs(np.mean([r["dll"] for r in h])) < 0.05
    assert abs(np.mean([r["carrier_hz"] for r in h]) - (if_hz + doppler)) < 15.0
