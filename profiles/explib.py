"""Experiment-only: run the Python binding against another build of libgnssacq.so (A/B timing, instrumented
builds).  The product loader (gnssacq/api.py) takes no override; this makes a throw-away copy of the package
(symlinks) whose libgnssacq.so is the requested file and puts it first on sys.path.  Call before importing gnssacq."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "assignment-for-aae6102_gnss-sdr_b200")


def use_lib(path=None):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not path:
        sys.path.insert(0, PKG)
        return
    src = os.path.join(PKG, "gnssacq")
    d = tempfile.mkdtemp(prefix="gnssacq_exp_")
    os.mkdir(os.path.join(d, "gnssacq"))
    for f in os.listdir(src):
        if f.endswith(".py"):
            os.symlink(os.path.join(src, f), os.path.join(d, "gnssacq", f))
    os.symlink(os.path.abspath(path), os.path.join(d, "gnssacq", "libgnssacq.so"))
    sys.path.insert(0, d)
