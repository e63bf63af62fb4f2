"""the Acquired_<name>_<skip>.mat hand-off file has the schema of the reference's own saved file."""
import json
import os
from types import SimpleNamespace

import numpy as np

import gnssacq
from gnssacq.matfile import FIELDS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_roundtrip_and_schema(tmp_path):
    g = json.load(open(os.path.join(GOLDEN, "acquired_opensky_5000.json")))
    acq = {f: np.array(g[f]) for f in FIELDS}
    file = SimpleNamespace(fileName="Opensky", skip=5000)
    assert gnssacq.acquired_filename(file) == "Acquired_Opensky_5000.mat"          # SDR_main.m:21
    path = gnssacq.save_acquired(acq, str(tmp_path / gnssacq.acquired_filename(file)))
    from scipy.io import loadmat
    m = loadmat(path, mat_dtype=True)["Acquired"]
    assert m.shape == (1, 1) and set(m.dtype.names) == set(FIELDS)
    for f in FIELDS:
        v = m[f][0, 0]
        assert v.dtype == np.float64 and v.shape == (1, 8)                         # 1xk double row vectors
    back = gnssacq.load_acquired(path)
    for f in FIELDS:
        assert np.array_equal(back[f], acq[f])


def test_empty_result_is_0x0(tmp_path):
    acq = {f: np.array([]) for f in FIELDS}
    path = gnssacq.save_acquired(acq, str(tmp_path / "Acquired_x_0.mat"))
    from scipy.io import loadmat
    m = loadmat(path, mat_dtype=True)["Acquired"]
    assert all(m[f][0, 0].size == 0 for f in FIELDS)                               # isempty(Acquired.sv), SDR_main.m:28
    assert gnssacq.load_acquired(path)["sv"].size == 0
