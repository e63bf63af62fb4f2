"""``initParameters.m`` restated (oracle; test infrastructure only).

Only the three structs ``acquisition.m`` reads are modelled -- ``file``,
``signal`` and ``acq`` (``SDR_MATLAB-main/initParameters.m:20-22,35-55``).
Field names are the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, List, Optional


@dataclass
class FileParams:
    """initParameters.m:20-22,35-38."""
    fileName: str = "Opensky"
    fid: Any = None              # open binary file object (the MATLAB file id)
    skip: int = 5000             # ms (:22)
    dataType: int = 2            # 1: real, 2: I/Q (:37)
    dataPrecision: int = 1       # 1: int8, 2: int16 (:38)


@dataclass
class SignalParams:
    """initParameters.m:41-47."""
    IF: float = 4.58e6
    Fs: float = 58e6
    Fc: float = 1575.42e6
    codeFreqBasis: float = 1.023e6
    ms: float = 1e-3
    Sample: int = 0
    codelength: float = 0.0

    def __post_init__(self) -> None:
        if not self.Sample:
            self.Sample = int(math.ceil(self.Fs * self.ms))           # :46
        if not self.codelength:
            self.codelength = self.codeFreqBasis * self.ms            # :47


@dataclass
class AcqParams:
    """initParameters.m:50-55 (``prnList`` is defined there but unused by acquisition.m:47)."""
    prnList: List[int] = field(default_factory=lambda: list(range(1, 33)))
    freqStep: float = 500.0
    freqMin: float = -10000.0
    freqNum: Optional[int] = None
    datalen: int = 20
    L: int = 10

    def __post_init__(self) -> None:
        if self.freqNum is None:
            self.freqNum = int(2 * abs(self.freqMin) / self.freqStep + 1)   # :53


def init_parameters(shape: str = "opensky"):
    """Return ``(file, signal, acq)`` with the reference defaults.

    ``shape="opensky"`` is initParameters.m verbatim (58 MHz, IF 4.58 MHz,
    int8 I/Q).  ``shape="urban"`` is the Urban recording's front end (26 MHz,
    IF 0, int8 I/Q): the repo only carries it as the ``%0`` alternative on
    initParameters.m:41 plus ``nAcquired_Urban_5000.mat``.
    """
    if shape == "opensky":
        return FileParams(), SignalParams(), AcqParams()
    if shape == "urban":
        return (FileParams(fileName="Urban"),
                SignalParams(IF=0.0, Fs=26e6),
                AcqParams())
    raise ValueError(shape)
