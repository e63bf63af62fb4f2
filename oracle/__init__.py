"""CPU oracle for the acquisition hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy float64 restatement of the reference receiver's
``acquisition.m`` (coarse search, lines 19-80; fine-frequency stage, lines
83-127), ``generateCAcode.m`` and the ``file/signal/acq`` parts of
``initParameters.m``.  It exists to *check* the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product
(``libgnssacq.so`` and the ``gnssacq`` Python package) never does, and has no
CPU fallback.

PARITY UNPINNED: the reference ships no tests, no golden input recording and
cannot be executed here (no MATLAB / Octave in the image).  The restatement is
therefore pinned by independent checks instead (see ``tests/test_oracle_*.py``):
the IS-GPS-200 first-ten-chip octal table for all 32 PRNs, brute-force
time-domain correlation in extended precision at random cells, truth recovery
on synthetic IF, and the output schema of the reference's saved
``Acquired_Opensky_5000.mat`` (committed as ``tests/golden/acquired_*.json``).
"""
from .params import init_parameters, FileParams, SignalParams, AcqParams  # noqa: F401
from .cacode import generate_ca_code  # noqa: F401
from .acquisition_ref import (  # noqa: F401
    acquisition,
    coarse_search,
    correlation_surface,
    fine_frequency,
    read_if_block,
    CoarseRow,
)
from . import tracking_ref  # noqa: F401  (trackingCT.m:75-150, the correlators' oracle)
from .synth import SynthSpec, SatSpec, synth_if, OPENSKY_TRUTH, URBAN_TRUTH  # noqa: F401
