"""Wall time of gnssacq_search (pageable host buffer in, rows out) for several builds of the library, same box:
python profiles/time_e2e.py name=path/to/lib.so ...   (median of 40 searches, Opensky- and Urban-shaped 32-PRN blocks)"""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, time
sys.path.insert(0, "/root/repo/profiles")
import explib
explib.use_lib(os.environ.get("AB_LIB"))
import numpy as np
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
for which in ("urban", "opensky"):
    spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
    raw = np.frombuffer(spec.read(0, 20), dtype=np.uint8).copy()
    with api.Searcher(gnssacq.make_config(fs_hz=fs, if_hz=if_hz)) as s:
        for _ in range(5):
            s.search(raw)
        wall, dev, h2d = [], [], []
        for _ in range(40):
            t0 = time.perf_counter()
            s.search(raw)
            wall.append((time.perf_counter() - t0) * 1e3)
            dev.append(s.last_stats.total_ms)
            h2d.append(s.last_stats.h2d_ms)
    med = lambda v: sorted(v)[len(v) // 2]
    print(which, "wall_ms %.3f" % med(wall), "device_total_ms %.3f" % med(dev), "h2d_ms %.3f" % med(h2d), "search_ms %.3f" % s.last_stats.search_ms, flush=True)
'''
for spec in sys.argv[1:]:
    name, path = spec.split("=", 1)
    for rnd in range(2):
        out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, AB_LIB=os.path.abspath(path)), capture_output=True, text=True)
        print(name, rnd, out.stdout.strip().replace("\n", " | ") if out.returncode == 0 else "FAILED " + out.stderr[-400:], flush=True)
