"""Host-side mirror of ``initParameters.m`` (the ``file``/``signal``/``acq`` structs).

Same field names and defaults as ``SDR_MATLAB-main/initParameters.m:20-22,35-55`` so code
written against the MATLAB structs reads the same here.
"""
from __future__ import annotations

import math
from types import SimpleNamespace


def initParameters(file_route: str | None = None, *, shape: str = "opensky"):
    """Return ``(file, signal, acq)``.  ``file.fid`` is opened when ``file_route`` is given."""
    file = SimpleNamespace(
        fileName="Opensky" if shape == "opensky" else "Urban",
        fileRoute=file_route,
        skip=5000,                 # ms            (initParameters.m:22)
        fid=open(file_route, "rb") if file_route else None,   # (:35)
        dataType=2,                # 1:I 2:IQ      (:37)
        dataPrecision=1,           # 1:int8 2:int16 (:38)
    )
    if shape == "opensky":
        IF, Fs = 4.58e6, 58e6      # (:41-42)
    elif shape == "urban":
        IF, Fs = 0.0, 26e6         # the `%0` alternative on :41; Urban front end
    else:
        raise ValueError(shape)
    signal = SimpleNamespace(IF=IF, Fs=Fs, Fc=1575.42e6, codeFreqBasis=1.023e6, ms=1e-3)
    signal.Sample = int(math.ceil(signal.Fs * signal.ms))          # (:46)
    signal.codelength = signal.codeFreqBasis * signal.ms           # (:47)
    acq = SimpleNamespace(prnList=list(range(1, 33)), freqStep=500.0, freqMin=-10000.0, datalen=20, L=10)
    acq.freqNum = int(2 * abs(acq.freqMin) / acq.freqStep + 1)     # (:53)
    return file, signal, acq
