"""Tiny end-to-end case for compute-sanitizer (one tool per gpurun call): every exchange mode at N = 6000."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200"]
import gnssacq
from gnssacq import api
from gnssacq.synth import Recording, Satellite
rec = Recording(fs=6e6, if_hz=1.25e6, samples_per_ms=6000, sats=[Satellite(3, 990.0, 1683, 1.2)])
raw = rec.read(0, 2)
long_raw = rec.read(0, 11)
for x in (1, 2, 3):
    cfg = gnssacq.make_config(fs_hz=6e6, if_hz=1.25e6, samples_per_ms=6000, freq_min_hz=-500.0, freq_step_hz=500.0,
                              freq_num=3, noncoh_blocks=2, prns=[3, 7], exchange=x)
    with api.Searcher(cfg) as s:
        rows = s.search(raw)
        ff = s.fine_frequency(long_raw, 10, [3], [rows[0].code_phase])
        print("exchange", x, [(r.prn, r.acquired, r.code_phase, r.doppler_bin) for r in rows], ff, flush=True)
