import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import numpy as np, gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
for name, spec, fs, if_hz in (("urban", urban_recording(), 26e6, 0.0), ("opensky", opensky_recording(), 58e6, 4.58e6)):
    raw = spec.read(0, 20)
    long_raw = spec.read(0, 11)
    cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz)
    with api.Searcher(cfg) as s:
        rows = s.search(raw)
        hit = [r for r in rows if r.acquired]
        prns, cps = [r.prn for r in hit], [r.code_phase for r in hit]
        for i in range(3):
            t0 = time.perf_counter(); ff = s.fine_frequency(long_raw, 10, prns, cps); dt = time.perf_counter() - t0
        truth = {x.prn: x.doppler_hz for x in spec.sats}
        print(name, "fine stage", len(prns), "SVs", round(dt * 1e3, 2), "ms wall (H2D 11 ms IF + tables + kernels + D2H);",
              "err Hz", [round(f - if_hz - truth.get(p, float('nan')), 1) for p, f in zip(prns, ff)], flush=True)
