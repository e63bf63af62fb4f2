"""How long is one block-iteration of a CTA group when it runs (almost) alone?  Rows = PRNs x bins are dealt
out whole (work_split=1), one row per group, so `rows` groups are busy for exactly K iterations each and the
rest exit at once: rows = 1 -> one CTA on 16 (8) SMs, rows = 9 (18) -> one CTA per SM, rows = 37 (74) -> the
full machine in one round.  search_ms / K = the latency chain of an iteration at that co-residency."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import numpy as np, gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
K = 20
for name, spec, fs, if_hz, per_sm1 in (("opensky", opensky_recording(), 58e6, 4.58e6, 9), ("urban", urban_recording(), 26e6, 0.0, 18)):
    raw = spec.read(0, K)
    for rows in (1, 2, per_sm1, 2 * per_sm1, 3 * per_sm1, 4 * per_sm1 + (1 if name == "opensky" else 2)):
        nprn = min(rows, 32)
        bins = (rows + nprn - 1) // nprn
        cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=list(range(1, nprn + 1)), freq_min_hz=0.0, freq_step_hz=500.0,
                                  freq_num=bins, noncoh_blocks=K, work_split=1)
        with api.Searcher(cfg) as s:
            best = 1e9
            for i in range(5):
                s.search(raw); st = s.last_stats
                best = min(best, st.search_ms)
            print(f"{name} rows={nprn * bins:3d} groups_resident={st.resident_clusters} search_ms={best:.4f} per_iteration_us={best / K * 1e3:.2f}"
                  f" cycles@1.965={best / K * 1.965e6:.0f}", flush=True)
