"""clock64 stamps of one steady-state iteration of CTA group 0 (experiment build with -DGNSS_TIMELINE):
per CTA rank and warp, cycles from the iteration start to the end of each phase of search_kernel_coop."""
import sys, ctypes as C
sys.path.insert(0, "/root/repo/profiles")
import explib
explib.use_lib(sys.argv[1] if len(sys.argv) > 1 else "/root/repo/profiles/r02/libgnssacq_tl.so")
import numpy as np, gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
NAMES = ["start", "p1 compute", "bar(p3 done)", "arrive", "p1 store", "bar", "p2", "spin", "bar", "p4", "p3"]
which = sys.argv[2] if len(sys.argv) > 2 else "opensky"
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1
spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
raw = spec.read(0, 20)
nprn = min(rows, 32); bins = (rows + nprn - 1) // nprn
cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=list(range(1, nprn + 1)), freq_min_hz=0.0, freq_num=bins, work_split=1)
with api.Searcher(cfg) as s:
    for i in range(3):
        s.search(raw)
    st = s.last_stats
    buf = np.zeros(16 * 16 * 32, dtype=np.uint64)
    assert api.lib.gnssacq_debug_timeline(buf.ctypes.data_as(C.c_void_p)) == 0
    R, W = st.cluster_ctas, st.threads // 32
    tl = buf[: R * W * 32].reshape(R, W, 32).astype(np.int64)
    print(f"{which} rows={nprn * bins} R={R} warps={W} search_ms={st.search_ms:.4f}")
    for r in (0, R // 2, R - 1):
        for w in range(W):
            t = tl[r, w, :11]
            print(f"rank {r:2d} warp {w}: " + "  ".join(f"{NAMES[i]}={t[i] - t[0]}" for i in range(1, 11)))
    # iteration length: stamp 0 of iteration 6 vs the same of ... only one iteration is stamped; report the phase ends
