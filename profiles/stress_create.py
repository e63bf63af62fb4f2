"""Race hunt at handle creation: create -> search -> destroy, many times; every handle must give the same bytes.
(r01: a synchronous cudaMemcpy from pageable memory followed by K0 on a non-blocking stream gave one wrong code
spectrum in ~1 of 100 handles; all create-time copies are on the handle's stream since.)"""
import sys
import time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for n, (spec, fs, if_hz, variants) in {
        26000: (urban_recording(), 26e6, 0.0, [(2, 512, 2), (0, 0, 0), (4, 256, 1)]),
        58000: (opensky_recording(), 58e6, 4.58e6, [(0, 0, 0), (4, 512, 2)])}.items():
    raw = spec.read(0, 2)
    for r, t, x in variants:
        first, bad, t0 = None, 0, time.time()
        for i in range(iters):
            cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, noncoh_blocks=2, cluster_ctas=r, threads=t, exchange=x)
            with api.Searcher(cfg) as s:
                got = [bytes(q) for q in s.search(raw)]
            if first is None:
                first = got
            elif got != first:
                bad += 1
                print("   MISMATCH handle", i, "prns", [j + 1 for j in range(32) if got[j] != first[j]], flush=True)
        print(n, (r, t, x), "handles", iters, "mismatching", bad, f"{time.time() - t0:.1f}s", flush=True)
