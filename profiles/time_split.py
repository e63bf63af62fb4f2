"""Search-kernel time against the number of PRNs of a handle (= one rank's shard at 32/n GPUs), block-granular
work split (default) against whole rows (work_split=1).  Usage: python profiles/time_split.py [urban|opensky]..."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording

for which in (sys.argv[1:] or ["urban", "opensky"]):
    spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
    raw = spec.read(0, 20)
    for n_prn in (1, 2, 4, 8, 16, 32):
        out = []
        for split in (1, 2):
            cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=range(1, n_prn + 1), work_split=split)
            with api.Searcher(cfg) as s:
                best = tot = 1e9
                for _ in range(6):
                    s.search(raw)
                    st = s.last_stats
                    best = min(best, st.search_ms)
                    tot = min(tot, st.total_ms)
                out.append((st.resident_clusters, round(best, 4), round(tot, 4)))
        print(which, "prns", n_prn, "rows(groups,search_ms,total_ms)", out[0], "blocks", out[1],
              "gain", round(out[0][1] / out[1][1], 3), flush=True)
