"""Race hunt: every engine variant, the same search repeated; any change in the result bytes is a race.
Usage: python profiles/stress_determinism.py [iters]"""
import sys
import time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
VARIANTS = {26000: [(2, 512, 1), (4, 256, 1), (4, 512, 1), (2, 512, 2), (4, 256, 2), (2, 512, 3), (4, 256, 3), (8, 128, 3)],
            58000: [(4, 512, 1), (8, 256, 1), (4, 512, 2), (4, 512, 3), (8, 256, 3), (16, 128, 3)]}
for n, (spec, fs, if_hz) in {26000: (urban_recording(), 26e6, 0.0), 58000: (opensky_recording(), 58e6, 4.58e6)}.items():
    raw = spec.read(0, 4)
    ref = None
    for r, t, x in VARIANTS[n]:
        for nprn in (32, 5):
            cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, noncoh_blocks=4, prns=range(1, nprn + 1),
                                      cluster_ctas=r, threads=t, exchange=x)
            t0 = time.time()
            with api.Searcher(cfg) as s:
                first = [bytes(q) for q in s.search(raw)]
                bad = 0
                for i in range(iters):
                    got = [bytes(q) for q in s.search(raw)]
                    if got != first:
                        bad += 1
                        if bad <= 3:
                            d = [j + 1 for j in range(len(got)) if got[j] != first[j]]
                            print("   MISMATCH iter", i, "prns", d, flush=True)
            print(n, (r, t, x), "prns", nprn, "iters", iters, "mismatching", bad, f"{time.time() - t0:.1f}s", flush=True)
