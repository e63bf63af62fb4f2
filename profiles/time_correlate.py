"""Latency of one gnssacq_correlate call (one integration period of a batch of channels, float64) against the
NumPy restatement of trackingCT.m:85-118 on one host core.  Usage: python profiles/time_correlate.py"""
import sys
import time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import numpy as np
import gnssacq
from gnssacq import api
from gnssacq.synth import opensky_recording
from oracle import tracking_ref as tr

fs, if_hz, n = 58e6, 4.58e6, 58000
raw = opensky_recording().read(0, 12)
x_all = tr.samples_of(raw, 2, 1)
EPL = [-0.5, 0.0, 0.5]
BANK25 = [round(0.6 - 0.05 * i, 2) for i in range(25)]
prns = [3, 4, 16, 22, 26, 27, 31, 32]
with api.Searcher(gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=[1])) as s:
    s.track_load(raw)
    for n_ch in (1, 8, 32):
        chans = [api.Channel(prn=prns[i % 8], num_samples=n, sample_offset=1000 * i + 17, carrier_hz=if_hz + 100.0 * i,
                             rem_phase=0.1 * i, code_hz=1.023e6 + 0.1 * i, rem_chip=0.01 * i) for i in range(n_ch)]
        for name, taps in (("E/P/L", EPL), ("25-tap bank", BANK25)):
            s.correlate(chans, taps)
            t0 = time.perf_counter()
            reps = 200
            for _ in range(reps):
                s.correlate(chans, taps)
            dt = (time.perf_counter() - t0) / reps
            t1 = time.perf_counter()
            c = chans[0]
            tr.correlate(x_all[c.sample_offset:c.sample_offset + n], fs, c.prn, c.carrier_hz, c.rem_phase, c.code_hz, c.rem_chip, taps)
            cpu = time.perf_counter() - t1
            print(f"{n_ch:2d} channels x {name:12s}: {dt * 1e6:7.1f} us per call ({n_ch / dt / 1e3:7.1f} k channel-ms/s, "
                  f"{n_ch * len(taps) * n / dt / 1e9:6.2f} G tap-samples/s) | oracle, one core: {cpu * 1e3:6.2f} ms per channel "
                  f"-> x{cpu * n_ch / dt:7.0f}", flush=True)
