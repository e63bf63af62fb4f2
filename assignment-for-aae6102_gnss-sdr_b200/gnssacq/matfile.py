"""Result hand-off files: the reference caches its stage outputs as MAT files next to the
script -- ``save(['Acquired_',file.fileName,'_',num2str(file.skip)],'Acquired')`` (SDR_main.m:23) -- and
reloads them instead of recomputing (SDR_main.m:21-27).  These helpers write / read exactly that file
from the Python twin, so a MATLAB user can consume GPU results without the MEX gateway.

Schema = the reference's own saved ``Acquired_Opensky_5000.mat``: a 1x1 struct ``Acquired`` with fields
``sv, SNR, Doppler, codedelay, fineFreq``, each a 1xk double row vector (0x0 when nothing is acquired).
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

FIELDS = ("sv", "SNR", "Doppler", "codedelay", "fineFreq")


def acquired_filename(file) -> str:
    """SDR_main.m:21,23 -- ``Acquired_<fileName>_<skip>.mat``."""
    return f"Acquired_{file.fileName}_{int(file.skip)}.mat"


def save_acquired(Acquired: Dict[str, np.ndarray], path: str, varname: str = "Acquired") -> str:
    from scipy.io import savemat
    rec = {}
    for f in FIELDS:
        v = np.asarray(Acquired[f], dtype=np.float64).reshape(1, -1)
        rec[f] = v if v.size else np.zeros((0, 0))
    savemat(path, {varname: rec}, format="5", oned_as="row")
    return path


def load_acquired(path: str, varname: str = "Acquired") -> Dict[str, np.ndarray]:
    from scipy.io import loadmat
    m = loadmat(path, mat_dtype=True)[varname]
    return {f: np.asarray(m[f][0, 0], dtype=np.float64).reshape(-1) for f in FIELDS}


def cached_acquisition(file, signal, acq, directory: str = ".", **kw):
    """The stage cache of SDR_main.m:21-27: load ``Acquired_<name>_<skip>.mat`` if present, else acquire and save."""
    from .acquisition import acquisition
    path = os.path.join(directory, acquired_filename(file))
    if os.path.exists(path):
        return load_acquired(path)
    out = acquisition(file, signal, acq, **kw)
    save_acquired(out, path)
    return out
