"""CPU emulation of the CUDA engine's pass functions (tests/emu/emu.cpp compiles the same GNSS_HD code
the kernels call): prime-factor maps, G layout, bin-shift identity, wipe-off and code
loaders, power accumulation.  Checked against NumPy.  Test double only -- never used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle.acquisition_ref import code_replica
from oracle.synth import SynthSpec, SatSpec, synth_if

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "assignment-for-aae6102_gnss-sdr_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libgnssemu.so")
    src = os.path.join(EMU_DIR, "emu.cpp")
    hdrs = [os.path.join(CSRC, h) for h in ("gnss_cplx.h", "gnss_radix.h", "gnss_engine.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, src, "-o", so])
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("Q,R", [(3, 1), (3, 2), (3, 4), (13, 2), (13, 4), (29, 4), (29, 8)])
def test_engine_passes_against_numpy(emu, Q, R):
    N = 2000 * Q
    fs = N * 1000.0
    if_hz = 4.58e6 if Q == 29 else 0.0
    sig = oracle.SignalParams(IF=if_hz, Fs=fs)
    sc = code_replica(sig, 5)
    outg = np.zeros(2 * N, np.float32)
    nat = np.zeros(2 * N, np.float32)
    assert emu.emu_code_spectrum(Q, R, P(sc.astype(np.int8)), P(outg)) == 0
    assert emu.emu_g_to_natural(Q, P(outg), P(nat)) == 0
    ref = np.conj(np.fft.fft(sc)) / N
    assert np.abs(nat.view(np.complex64) - ref).max() <= 2e-6 * np.abs(ref).max()

    spec = SynthSpec(fs=fs, if_hz=if_hz, samples_per_ms=N, sats=[SatSpec(5, 1234.0, 777, 3.0)])
    K = 2
    raw = np.frombuffer(synth_if(spec, 0, K), np.int8)
    xs = raw[0::2].astype(float) + 1j * raw[1::2].astype(float)
    n1 = np.arange(1, N + 1)
    f0 = if_hz - 1000.0
    NX = 16 * (2 * Q - 1) * 125
    xg = np.zeros((K, 2 * NX), np.float32)
    for k in range(K):
        blk = np.ascontiguousarray(raw[2 * N * k:2 * N * (k + 1)])
        assert emu.emu_wipe_spectrum(Q, R, P(blk), 2, 1, 1, C.c_double(f0), C.c_double(fs),
                                     C.c_float(0), C.c_float(0), P(xg[k])) == 0
        assert emu.emu_gx_to_natural(Q, P(xg[k]), P(nat)) == 0
        xr = np.fft.fft(xs[N * k:N * (k + 1)] * np.exp(2j * np.pi * f0 * n1 / fs))
        assert np.abs(nat.view(np.complex64) - xr).max() <= 2e-6 * np.abs(xr).max()

    cfft = np.fft.fft(sc)
    for s in (0, 1, -3, 7):
        acc = np.zeros(N, np.float32)
        assert emu.emu_search_row(Q, R, P(outg), P(xg), K, s, P(acc)) == 0
        f = f0 + s * 1000.0
        want = np.zeros(N)
        for k in range(K):
            t1 = xs[N * k:N * (k + 1)] * np.exp(2j * np.pi * f * n1 / fs)
            want += np.abs(np.fft.ifft(cfft * np.conj(np.fft.fft(t1)))) ** 2
        assert np.abs(acc - want).max() <= 5e-6 * want.max()
        assert int(acc.argmax()) == int(want.argmax())
        acc2 = np.zeros(N, np.float32)           # same row through the transposed (XT) exchange path
        assert emu.emu_search_row_xt(Q, R, P(outg), P(xg), K, s, P(acc2)) == 0
        assert np.array_equal(acc2, acc)


def test_fine_frequency_loader_against_numpy(emu):
    """FineLoader + natural store: one decimated, modulated N-point transform of the code-stripped long block."""
    Q, R, L, K = 3, 2, 10, 4
    N = 2000 * Q
    fs = N * 1000.0
    sig = oracle.SignalParams(IF=1.25e6, Fs=fs)
    spec = SynthSpec(fs=fs, if_hz=1.25e6, samples_per_ms=N, sats=[SatSpec(5, 1234.0, 777, 3.0)])
    raw = np.frombuffer(synth_if(spec, 0, L + 1), np.int8)
    xs = raw[0::2].astype(float) + 1j * raw[1::2].astype(float)
    t = np.arange(1, L * N + 1, dtype=np.float64)
    chip = np.fmod(np.floor((1.0 / fs * t) / (1.0 / sig.codeFreqBasis)), 1023.0).astype(np.uint16)
    ca = oracle.generate_ca_code(5).astype(np.int8)
    cd = 777
    start = N - cd - 1
    F = L * N * K
    s_long = xs[start:start + L * N] * ca[chip]
    for n2, r in ((0, 0), (3, 1), (9, 3)):
        out = np.zeros(2 * N, np.float32)
        rc = emu.emu_fine_unit(Q, R, P(raw), P(chip), P(ca), 2, 1, C.c_float(0), C.c_float(0), start, L, n2, r,
                               C.c_longlong(F), P(out))
        assert rc == 0
        n = L * np.arange(N) + n2
        want = np.fft.fft(s_long[n] * np.exp(-2j * np.pi * ((n * r) % F) / F))
        assert np.abs(out.view(np.complex64) - want).max() <= 3e-6 * np.abs(want).max()
