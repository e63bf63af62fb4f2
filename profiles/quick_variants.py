import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
which = sys.argv[1]
variants = [tuple(int(x) for x in v.split(",")) for v in sys.argv[2:]]
spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
raw = spec.read(0, 20)
for r, t, x in variants:
    cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, cluster_ctas=r, threads=t, exchange=x)
    try:
        with api.Searcher(cfg) as s:
            best = 1e9
            for i in range(5):
                rows = s.search(raw); st = s.last_stats; best = min(best, st.search_ms)
            print(which, r, t, x, "clusters", st.resident_clusters, "search_ms", round(best, 3), "acq", [q.prn for q in rows if q.acquired], flush=True)
    except Exception as e:
        print(which, r, t, x, "FAILED", e, flush=True)
