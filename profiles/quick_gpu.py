import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import numpy as np, gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
for name, spec, fs, if_hz, variants in (("urban", urban_recording(), 26e6, 0.0, [(2,512,1),(4,256,1),(2,512,2),(4,256,2),(4,512,2)]), ("opensky", opensky_recording(), 58e6, 4.58e6, [(4,512,1),(4,512,2),(8,256,2)])):
    raw = spec.read(0, 20)
    for r,t,x in variants:
        cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, cluster_ctas=r, threads=t, exchange=x)
        try:
            with api.Searcher(cfg) as s:
                for i in range(3):
                    rows = s.search(raw); st = s.last_stats
                print(name, r, t, x, "clusters", st.resident_clusters, "search_ms", round(st.search_ms,3), "k1", round(st.wipeoff_fft_ms,3), "total", round(st.total_ms,3), "acq", [x.prn for x in rows if x.acquired], flush=True)
        except Exception as e:
            print(name, r, t, x, "FAILED", e, flush=True)
