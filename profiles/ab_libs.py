"""Same-box A/B of two builds of libgnssacq.so (boxes differ by several per cent, so only same-run numbers
compare).  Usage: python profiles/ab_libs.py name=path/to/lib.so name2=path2 [--prns 4,32] [--rounds 3]"""
import os
import subprocess
import sys

CHILD = r'''
import os
import sys
sys.path.insert(0, "/root/repo/profiles")
import explib
explib.use_lib(os.environ.get("AB_LIB"))
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording
prns = [int(x) for x in sys.argv[1].split(",")]
for which in ("urban", "opensky"):
    spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
    raw = spec.read(0, 20)
    for n in prns:
        with api.Searcher(gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=range(1, n + 1), **({'work_split': int(os.environ['AB_WORK_SPLIT'])} if os.environ.get('AB_WORK_SPLIT') else {}), **({'threads': int(os.environ['AB_THREADS_' + which.upper()].split('x')[-1]), 'cluster_ctas': int(os.environ['AB_THREADS_' + which.upper()].split('x')[0]) if 'x' in os.environ['AB_THREADS_' + which.upper()] else 0} if os.environ.get('AB_THREADS_' + which.upper()) else {}))) as s:
            best = 1e9
            for _ in range(8):
                rows = s.search(raw)
                best = min(best, s.last_stats.search_ms)
            import hashlib, ctypes
            digest = hashlib.sha1(b"".join(ctypes.string_at(ctypes.addressof(r), ctypes.sizeof(r)) for r in rows)).hexdigest()[:10]
        print(which, n, round(best, 4), digest, flush=True)
'''

libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
prns = "4,32"
rounds = 3
for i, a in enumerate(sys.argv):
    if a == "--prns":
        prns = sys.argv[i + 1]
    if a == "--rounds":
        rounds = int(sys.argv[i + 1])
best = {}
digests = {}
for r in range(rounds):
    for name, path in libs:
        env = dict(os.environ, AB_LIB=os.path.abspath(path.split("@")[0]))
        if "@" in path:                                  # name=lib.so@1 -> work_split=1 ; name=lib.so@@160 -> threads=160 (opensky)
            parts = path.split("@")
            if parts[1]:
                env["AB_WORK_SPLIT"] = parts[1]
            if len(parts) > 2 and parts[2]:
                env["AB_THREADS_OPENSKY"] = parts[2]
            if len(parts) > 3 and parts[3]:          # name=lib.so@@160@160 -> threads for opensky, urban
                env["AB_THREADS_URBAN"] = parts[3]
        out = subprocess.run([sys.executable, "-c", CHILD, prns], env=env, capture_output=True, text=True)
        if out.returncode:
            print(name, "FAILED", out.stderr[-500:])
            continue
        for line in out.stdout.split("\n"):
            if line.strip():
                which, n, ms, digest = line.split()
                key = (which, int(n), name)
                best[key] = min(best.get(key, 1e9), float(ms))
                digests[key] = digest          # result rows of the last search: equal digests = byte-identical tables
for which in ("urban", "opensky"):
    for n in [int(x) for x in prns.split(",")]:
        print(which, "prns", n, {name: best.get((which, n, name)) for name, _ in libs},
              "rows", {name: digests.get((which, n, name)) for name, _ in libs}, flush=True)
