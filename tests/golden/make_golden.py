"""Regenerates the fixtures in this directory.  Run from the repo root, in the build container
(needs /root/reference for the first two files; the GPU box never runs this).

  acquired_opensky_5000.json / nacquired_urban_5000.json : the reference's own saved acquisition results
      (SDR_MATLAB-main/Acquired_Opensky_5000.mat, nAcquired_Urban_5000.mat), values only.
  small_rows.json : oracle rows on a seeded synthetic input, to catch oracle drift.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))


def from_mat(path, var, if_hz, out):
    import scipy.io as sio
    m = sio.loadmat(path, mat_dtype=True)[var]
    fields = list(m.dtype.names)
    d = {"source": os.path.relpath(path, "/root/reference"), "fields": fields, "IF": if_hz}
    for f in fields:
        v = m[f][0, 0]
        assert v.dtype == np.float64 and v.shape[0] == 1
        d[f] = [float(x) for x in v.ravel()]
    json.dump(d, open(os.path.join(HERE, out), "w"), indent=1)


def small_rows():
    from helpers import structs, small_spec, oracle_rows
    from oracle.synth import synth_if
    fs, if_hz, datalen, seed, prns = 6e6, 1.25e6, 2, 6102, [1, 3, 7, 22, 30]
    file, signal, acq = structs(fs, if_hz, datalen=datalen)
    raw = synth_if(small_spec(fs, if_hz, int(signal.Sample), seed=seed), 0, datalen)
    rows = oracle_rows(raw, file, signal, acq, prns)
    d = {"fs": fs, "if": if_hz, "datalen": datalen, "seed": seed, "prns": prns,
         "rows": [dict(prn=r.prn, code_phase=r.code_phase, doppler_bin=r.doppler_bin, acquired=r.acquired,
                       peak=r.peak, snr_db=r.snr_db) for r in rows]}
    json.dump(d, open(os.path.join(HERE, "small_rows.json"), "w"), indent=1)


if __name__ == "__main__":
    ref = "/root/reference/SDR_MATLAB-main"
    if os.path.isdir(ref):
        from_mat(f"{ref}/Acquired_Opensky_5000.mat", "Acquired", 4.58e6, "acquired_opensky_5000.json")
        from_mat(f"{ref}/nAcquired_Urban_5000.mat", "nAcquired", 0.0, "nacquired_urban_5000.json")
    small_rows()
