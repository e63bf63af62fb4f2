"""One configurable search for ncu captures: python profiles/one_search.py <urban|opensky> <n_prn> <work_split> [reps]"""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
which, n_prn, ws = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
raw = spec.read(0, 20)
with api.Searcher(gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=range(1, n_prn + 1), work_split=ws)) as s:
    for _ in range(reps):
        s.search(raw)
    print(which, n_prn, ws, "search_ms", round(s.last_stats.search_ms, 4), "schedule", s.last_stats.work_split)
