"""Device-side tracking loop (gnssacq_track) against the host-driven loop over gnssacq_correlate and the NumPy
restatement: wall time per tracked millisecond.  Usage: python profiles/time_track.py"""
import sys
import time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import numpy as np
import gnssacq
from gnssacq import api
from gnssacq.synth import opensky_recording
from oracle import tracking_ref as tr

fs, if_hz, n, periods = 58e6, 4.58e6, 58000, 1000
rec = opensky_recording()
raw = rec.read(0, periods + 3)
truth = list(zip((3, 4, 16, 22, 26, 27, 31, 32), (990.0, -3095.0, -305.0, 1565.0, 1835.0, -3225.0, 1045.0, 3345.0),
                 (3683, 12701, 26051, 2610, 57908, 49778, 39064, 20170)))     # gnssacq/synth.py::opensky_recording
start = [api.Channel(prn=t[0], num_samples=0, sample_offset=n - t[2] + 1, carrier_hz=if_hz + t[1] + 5.0, rem_phase=0.0,
                     code_hz=1.023e6, rem_chip=0.0) for t in truth]
with api.Searcher(gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=[1])) as s:
    s.track_load(raw)
    for n_ch in (1, 8):
        s.track(start[:n_ch], periods)                      # (also sizes the record buffer)
        t0 = time.perf_counter()
        recs = s.track(start[:n_ch], periods)
        dt = time.perf_counter() - t0
        lock = [float(np.median([np.hypot(r.P_i, r.P_q) for r in ch[50:]])) for ch in recs]
        print(f"gnssacq_track      : {n_ch} channels x {periods} ms in {dt * 1e3:7.2f} ms = {dt / periods * 1e6:6.1f} us per ms "
              f"({periods * 1e-3 / dt:5.1f} x real time); median |P| {np.round(lock, 0)}", flush=True)
        # host-driven: one gnssacq_correlate call per ms, loop filters in Python (oracle module's close_loops)
        sts = [tr.ChannelState(prn=c.prn, carrier_basis_hz=c.carrier_hz, carrier_hz=c.carrier_hz, sample_pos=c.sample_offset)
               for c in start[:n_ch]]
        t0 = time.perf_counter()
        for ms in range(100):
            ns = [tr.num_samples(st.code_hz, fs, st.rem_chip) for st in sts]
            chans = [api.Channel(prn=st.prn, num_samples=k, sample_offset=st.sample_pos, carrier_hz=st.carrier_hz,
                                 rem_phase=st.rem_phase, code_hz=st.code_hz, rem_chip=st.rem_chip) for st, k in zip(sts, ns)]
            gi, gq = s.correlate(chans, [-0.5, 0.0, 0.5])
            for c, st in enumerate(sts):
                tr.close_loops(st, gi[c], gq[c], ns[c], fs)
        dh = (time.perf_counter() - t0) / 100
        print(f"correlate per ms   : {n_ch} channels: {dh * 1e6:6.1f} us per ms (Python loop filters included)", flush=True)
x_all = tr.samples_of(raw[:2 * n * 12], 2, 1)
st = tr.ChannelState(prn=start[0].prn, carrier_basis_hz=start[0].carrier_hz, carrier_hz=start[0].carrier_hz, sample_pos=start[0].sample_offset)
t0 = time.perf_counter()
for ms in range(8):
    k = tr.num_samples(st.code_hz, fs, st.rem_chip)
    i, q = tr.correlate(x_all[st.sample_pos:st.sample_pos + k], fs, st.prn, st.carrier_hz, st.rem_phase, st.code_hz, st.rem_chip, [-0.5, 0.0, 0.5])
    tr.close_loops(st, i, q, k, fs)
print(f"oracle (NumPy)     : 1 channel: {(time.perf_counter() - t0) / 8 * 1e3:6.2f} ms per ms on one core", flush=True)
