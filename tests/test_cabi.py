"""C-ABI surface checks that need no GPU: the library loads, exports every symbol include/gnssacq.h
declares, struct layouts agree, host-side tables match the oracle, and errors are reported (not raised
across the ABI).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from oracle.acquisition_ref import code_replica
import gnssacq
from gnssacq import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gnssacq.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnssacq_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(api.lib, n), n
    assert set(names) == set(api.EXPORTS)


def test_struct_layout_matches_header():
    # compile a tiny C probe against the header and compare sizeof/offsets with the ctypes mirror
    import subprocess, tempfile
    probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "gnssacq.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(gnssacq_loop_params), sizeof(gnssacq_track_record),
   offsetof(gnssacq_track_record, num_samples), sizeof(gnssacq_config), offsetof(gnssacq_config, prn),
   offsetof(gnssacq_config, snr_threshold_db), offsetof(gnssacq_config, keep_surface),
   sizeof(gnssacq_result), sizeof(gnssacq_stats), sizeof(gnssacq_channel), offsetof(gnssacq_channel, carrier_hz));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    want = [C.sizeof(api.LoopParams), C.sizeof(api.TrackRecord), api.TrackRecord.num_samples.offset, C.sizeof(api.Config), api.Config.prn.offset, api.Config.snr_threshold_db.offset,
            api.Config.keep_surface.offset, C.sizeof(api.Result), C.sizeof(api.Stats), C.sizeof(api.Channel),
            api.Channel.carrier_hz.offset]
    assert got == want


def test_default_config_is_init_parameters():
    cfg = gnssacq.default_config()
    f, s, a = oracle.init_parameters("opensky")
    assert (cfg.fs_hz, cfg.if_hz, cfg.code_hz, cfg.samples_per_ms) == (s.Fs, s.IF, s.codeFreqBasis, s.Sample)
    assert (cfg.data_type, cfg.data_precision) == (f.dataType, f.dataPrecision)
    assert (cfg.freq_min_hz, cfg.freq_step_hz, cfg.freq_num, cfg.noncoh_blocks) == (a.freqMin, a.freqStep, a.freqNum, a.datalen)
    assert cfg.n_prn == 32 and list(cfg.prn[:32]) == list(range(1, 33)) and cfg.snr_threshold_db == 12.0
    assert api.lib.gnssacq_if_bytes(C.byref(cfg)) == 58000 * 2 * 20


def test_ca_code_and_replica_match_oracle():
    for prn in range(1, 52):
        assert np.array_equal(api.ca_code(prn), oracle.generate_ca_code(prn).astype(np.int8)), prn
    for shape in ("opensky", "urban"):
        _, s, _ = oracle.init_parameters(shape)
        cfg = gnssacq.make_config(fs_hz=s.Fs, if_hz=s.IF, samples_per_ms=s.Sample)
        for prn in (1, 17, 32):
            assert np.array_equal(api.code_replica(cfg, prn), code_replica(s, prn).astype(np.int8))
    assert api.lib.gnssacq_ca_code(0, np.zeros(1023, np.int8).ctypes.data) == -1
    assert api.lib.gnssacq_ca_code(52, np.zeros(1023, np.int8).ctypes.data) == -1


@pytest.mark.parametrize("kw,code", [
    (dict(samples_per_ms=58001), -2), (dict(samples_per_ms=34000, fs_hz=34e6), -2),
    (dict(data_type=3), -1), (dict(data_precision=2, data_type=1), -1), (dict(freq_num=0), -1),
    (dict(noncoh_blocks=0), -1), (dict(prns=[0]), -1), (dict(prns=[1, 99]), -1), (dict(coh_ms=0), -1),
])
def test_create_rejects_bad_config(kw, code):
    cfg = gnssacq.make_config(**kw)
    h = C.c_void_p()
    rc = api.lib.gnssacq_create(C.byref(cfg), C.byref(h))
    assert rc == code and not h.value
    assert api.lib.gnssacq_last_error(None)


def test_null_arguments_are_errors_not_crashes():
    assert api.lib.gnssacq_create(None, None) == -1
    assert api.lib.gnssacq_destroy(None) == -1
    assert api.lib.gnssacq_config_default(None) == -1
    assert api.lib.gnssacq_if_bytes(None) == 0
    assert api.lib.gnssacq_search(None, None, 0, None, None) == -1
    assert api.lib.gnssacq_sweep(None, None, 0, 0, None, None) == -1
    assert api.lib.gnssacq_track_load(None, None, 0) == -1
    assert api.lib.gnssacq_correlate(None, 0, None, 0, None, None, None) == -1
    assert api.lib.gnssacq_track(None, 0, None, None, 0, None) == -1
    assert api.lib.gnssacq_loop_params_default(None) == -1
    assert api.lib.gnssacq_sweep_file(None, b"/nonexistent", 0, 100, 1, None, None) == -1
    assert api.lib.gnssacq_fetch_results(None, None, None) == -1
    # the multi-GPU exchange: no handle -> state / argument errors, never a crash
    assert api.lib.gnssacq_shard_plan(None, 0, 1, None, None) == -1
    assert api.lib.gnssacq_xchg_root(None, None, None) == -1
    assert api.lib.gnssacq_xchg_attach(None, None, None) == -1
    assert api.lib.gnssacq_xchg_attach_local(None, None, None) == -1
    assert not api.lib.gnssacq_xchg_if_buffer(None)
    assert api.lib.gnssacq_xchg_enqueue(None, None, 0) == -7
    assert api.lib.gnssacq_xchg_finish(None) == -7
    assert api.lib.gnssacq_xchg_fetch(None, None, None) == -7
    full = gnssacq.make_config(prns=[1, 2, 3])
    mine, sh = api.Config(), api.Shard()
    assert api.lib.gnssacq_shard_plan(C.byref(full), 3, 3, C.byref(mine), C.byref(sh)) == -1      # rank out of range
    assert api.lib.gnssacq_shard_plan(C.byref(full), 0, 63, C.byref(mine), C.byref(sh)) == -1     # more shards than flag words
    full.bin_first, full.bin_count = 2, 5
    assert api.lib.gnssacq_shard_plan(C.byref(full), 0, 2, C.byref(mine), C.byref(sh)) == -1      # wants the full grid
    bad = gnssacq.make_config(prns=[1], bin_first=40, bin_count=5)                                  # 41-bin grid
    h = C.c_void_p()
    assert api.lib.gnssacq_create(C.byref(bad), C.byref(h)) == -1 and not h.value
    lp = api.LoopParams()
    assert api.lib.gnssacq_loop_params_default(C.byref(lp)) == 0
    assert (lp.dll_bw, lp.dll_damp, lp.dll_gain, lp.pll_bw, lp.pll_damp, lp.pll_gain, lp.spacing_chips) == (2.0, 0.707, 0.1, 15.0, 0.707, 0.25, 0.5)   # initParameters.m:59-65


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gnssacq.GnssAcqError) as e:
        gnssacq.Searcher(gnssacq.default_config())
    assert e.value.code == -5


def test_mex_gateway_compiles_against_header():
    """matlab/gnssacq_mex.c is source-only here (no MATLAB): syntax/type-check it against the real
    gnssacq.h and a stub mex.h so it cannot drift from the ABI."""
    import subprocess
    src = os.path.join(ROOT, "assignment-for-aae6102_gnss-sdr_b200", "matlab", "gnssacq_mex.c")
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "stubs"), src])


def test_product_and_oracle_generators_agree():
    """gnssacq/synth.py (product side, library C/A tables) and oracle/synth.py (NumPy only) are independent
    implementations of SURVEY Appendix C: same bytes for the same recording window."""
    from gnssacq.synth import urban_recording, opensky_recording
    from oracle.synth import urban_spec, opensky_spec, synth_if
    assert urban_recording().read(41, 2) == synth_if(urban_spec(), 41, 2)
    assert opensky_recording(seed=7).read(5, 1) == synth_if(opensky_spec(seed=7), 5, 1)


def test_c_example_links_against_the_library(tmp_path):
    """examples/acquire.c uses the ABI from plain C (no Python, no torch): it must compile and link."""
    import subprocess
    pkg = os.path.join(ROOT, "assignment-for-aae6102_gnss-sdr_b200", "gnssacq")
    exe = str(tmp_path / "acquire")
    subprocess.check_call(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "acquire.c"), "-L", pkg, "-lgnssacq", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=pkg))
    assert r.returncode == 2 and "usage" in r.stderr
