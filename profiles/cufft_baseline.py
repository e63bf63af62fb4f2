"""Library baseline for the record (NOT product, NOT a test oracle): the same acquisition written the obvious way on
torch.fft (cuFFT), complex64, everything resident in HBM, timed with CUDA events next to libgnssacq on the same
box.  Forward spectra are shared by all PRNs (the same algorithmic saving the product takes), the code spectra
are cached, each PRN is one batched inverse transform of B*K rows.  Usage: python profiles/cufft_baseline.py"""
import sys
import time
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import numpy as np
import torch
import gnssacq
from gnssacq import api
from gnssacq.synth import urban_recording, opensky_recording

dev = torch.device("cuda:0")
for which in ("urban", "opensky"):
    spec, fs, if_hz = (urban_recording(), 26e6, 0.0) if which == "urban" else (opensky_recording(), 58e6, 4.58e6)
    N, K, B, P = int(fs * 1e-3), 20, 41, 32
    raw = spec.read(0, K)
    cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz)
    with api.Searcher(cfg) as s:
        best = 1e9
        for _ in range(5):
            rows = s.search(raw)
            best = min(best, s.last_stats.wipeoff_fft_ms + s.last_stats.search_ms + s.last_stats.finalize_ms)
    ours = [(r.code_phase, r.doppler_bin) for r in rows]

    iq = torch.from_numpy(np.frombuffer(raw, np.int8).astype(np.float32)).to(dev)
    x = torch.complex(iq[0::2], iq[1::2]).reshape(K, N)
    n = torch.arange(1, N + 1, device=dev, dtype=torch.float64)
    freqs = if_hz + (-10000.0 + 500.0 * torch.arange(B, device=dev, dtype=torch.float64))
    carrier = torch.exp(1j * 2 * np.pi * freqs[:, None] * n[None, :] / fs).to(torch.complex64)          # [B, N]
    codes = np.stack([api.code_replica(cfg, p).astype(np.float32) for p in range(1, P + 1)])
    cfft = torch.fft.fft(torch.from_numpy(codes).to(dev).to(torch.complex64), dim=1)                    # cached, [P, N]

    def run():
        spectra = torch.conj(torch.fft.fft(x[None, :, :] * carrier[:, None, :], dim=2))                 # [B, K, N]
        out = []
        for p in range(P):
            corr = torch.fft.ifft(spectra * cfft[p][None, None, :], dim=2)                              # [B, K, N]
            acc = (corr.real ** 2 + corr.imag ** 2).sum(dim=1)                                          # [B, N]
            flat = torch.argmax(acc)
            out.append(flat)
        return torch.stack(out)

    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(3):
        e0.record(); idx = run(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    idx = idx.cpu().numpy()
    theirs = [(int(i % N), int(i // N)) for i in idx]
    agree = sum(a == b for a, b in zip(ours, theirs))
    cells = P * B * N
    print(f"{which}: libgnssacq kernels {best:.3f} ms ({cells / best / 1e6:.2f} G cells/s) | torch.fft/cuFFT pipeline "
          f"{min(times):.2f} ms ({cells / min(times) / 1e6:.2f} G cells/s) | ratio {min(times) / best:.1f}x | "
          f"same (code phase, bin) for {agree}/32 PRNs", flush=True)
