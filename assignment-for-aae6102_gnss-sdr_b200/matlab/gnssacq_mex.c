/* gnssacq_mex.c -- MEX gateway from MATLAB to libgnssacq.so (include/gnssacq.h).
 *
 *   rows     = gnssacq_mex(raw_int8_or_int16, cfg)                        coarse search
 *   fineFreq = gnssacq_mex(longraw, cfg, L, sv, codedelay)                 fine-frequency stage
 *              gnssacq_mex(segment, cfg, 'track_load')                      recording segment -> HBM (tracking)
 *   [I, Q]   = gnssacq_mex(channels, cfg, spacing, 'correlate')            one integration period, all channels
 *              channels: 7 x n double, rows [prn; numSample; sample_offset; carrierFreq; remPhase; codeFreq;
 *              remChip] (trackingCT.m:42-58,78); I, Q: n x numel(spacing) (trackingCT.m:115-117)
 *   rec      = gnssacq_mex(channels, cfg, loops, n_periods, 'track')       closed DLL/PLL loop on the device
 *              loops: [DLLBW DLLDamp DLLGain PLLBW PLLDamp PLLGain CorrelatorSpacing]; rec: 14 x n_periods x n,
 *              rows [P_i P_q E_i E_q L_i L_q PLLdiscri DLLdiscri remChip codeFreq carrierFreq remPhase
 *              sample_end numSample] (trackingCT.m:153-172)
 *
 * `raw` is the block acquisition.m:29/34 reads, passed as int8 (or int16) WITHOUT conversion to
 * double; `cfg` is a scalar struct whose fields are named after gnssacq_config.  Returns an
 * n_prn x 8 double matrix, one row per searched PRN:
 *   [prn acquired code_phase doppler_bin doppler_hz peak noise_meansq snr_db]
 * One handle is kept in a static between calls (mexLock) and rebuilt only when cfg changes; it is
 * destroyed by mexAtExit.  Library errors become MATLAB errors gnssacq:<code>.
 *
 * SOURCE-ONLY DELIVERABLE: neither MATLAB nor Octave (mex.h) exists in the build image, so this
 * file is compile-checked against tests/stubs/mex.h only.  Build on a MATLAB host with
 *   mex -I<repo>/include gnssacq_mex.c -L<repo>/assignment-for-aae6102_gnss-sdr_b200/gnssacq -lgnssacq
 */
#include <stdint.h>
#include <string.h>
#include "mex.h"
#include "gnssacq.h"

static gnssacq_handle* g_handle = NULL;
static gnssacq_config g_cfg;
static int g_have_cfg = 0;

static void at_exit(void) {
    if (g_handle) { gnssacq_destroy(g_handle); g_handle = NULL; }
}

static double field(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxGetField(s, 0, name);
    return (f && mxGetNumberOfElements(f) >= 1) ? mxGetScalar(f) : dflt;
}

static void fail(int rc, const char* msg) {
    char id[32];
    sprintf(id, "gnssacq:e%d", -rc);
    mexErrMsgIdAndTxt(id, "%s", msg ? msg : "gnssacq error");
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    gnssacq_config c;
    const mxArray* prn;
    gnssacq_result* rows;
    size_t nbytes;
    double* out;
    int rc, i, track_mode;

    if (nrhs < 2 || nrhs > 5 || !mxIsStruct(prhs[1]))
        fail(GNSSACQ_ERR_INVALID_ARG, "usage: rows = gnssacq_mex(raw, cfg) | fineFreq = gnssacq_mex(longraw, cfg, L, sv, codedelay) | "
                                      "gnssacq_mex(segment, cfg, 'track_load') | [I, Q] = gnssacq_mex(channels, cfg, spacing, 'correlate')");
    track_mode = (nrhs == 5 && mxIsChar(prhs[4]));
    if (nrhs != 4 && !track_mode && !(mxIsInt8(prhs[0]) || mxIsInt16(prhs[0]))) fail(GNSSACQ_ERR_INVALID_ARG, "raw must be int8 or int16");

    gnssacq_config_default(&c);
    c.fs_hz = field(prhs[1], "fs_hz", c.fs_hz);
    c.if_hz = field(prhs[1], "if_hz", c.if_hz);
    c.code_hz = field(prhs[1], "code_hz", c.code_hz);
    c.samples_per_ms = (int32_t)field(prhs[1], "samples_per_ms", c.samples_per_ms);
    c.data_type = (int32_t)field(prhs[1], "data_type", c.data_type);
    c.data_precision = (int32_t)field(prhs[1], "data_precision", c.data_precision);
    c.freq_min_hz = field(prhs[1], "freq_min_hz", c.freq_min_hz);
    c.freq_step_hz = field(prhs[1], "freq_step_hz", c.freq_step_hz);
    c.freq_num = (int32_t)field(prhs[1], "freq_num", c.freq_num);
    c.noncoh_blocks = (int32_t)field(prhs[1], "noncoh_blocks", c.noncoh_blocks);
    c.coh_ms = (int32_t)field(prhs[1], "coh_ms", c.coh_ms);
    c.snr_threshold_db = field(prhs[1], "snr_threshold_db", c.snr_threshold_db);
    c.device = (int32_t)field(prhs[1], "device", -1);
    prn = mxGetField(prhs[1], 0, "prn");
    if (prn) {
        size_t n = mxGetNumberOfElements(prn);
        const double* p = mxGetPr(prn);
        if (n > GNSSACQ_MAX_PRN) fail(GNSSACQ_ERR_INVALID_ARG, "too many PRNs");
        c.n_prn = (int32_t)n;
        for (i = 0; i < GNSSACQ_MAX_PRN; ++i) c.prn[i] = (i < (int)n) ? (int32_t)p[i] : 0;
    }

    if (!g_handle || !g_have_cfg || memcmp(&c, &g_cfg, sizeof c) != 0) {
        if (g_handle) { gnssacq_destroy(g_handle); g_handle = NULL; }
        rc = gnssacq_create(&c, &g_handle);
        if (rc != GNSSACQ_OK) fail(rc, gnssacq_last_error(NULL));
        if (!g_have_cfg) { mexLock(); mexAtExit(at_exit); }
        g_cfg = c;
        g_have_cfg = 1;
    }

    nbytes = mxGetNumberOfElements(prhs[0]) * mxGetElementSize(prhs[0]);
    if (nrhs == 3) {                                   /* tracking: keep the segment in HBM */
        rc = gnssacq_track_load(g_handle, mxGetData(prhs[0]), nbytes);
        if (rc != GNSSACQ_OK) fail(rc, gnssacq_last_error(g_handle));
        return;
    }
    if (nrhs == 4) {                                   /* tracking correlators (trackingCT.m:85-118) */
        const double* m = mxGetPr(prhs[0]);
        const int n_ch = (int)mxGetN(prhs[0]), n_taps = (int)mxGetNumberOfElements(prhs[2]);
        gnssacq_channel* ch;
        double *oi, *oq, *pi, *pq;
        int t;
        if (!mxIsDouble(prhs[0]) || mxGetM(prhs[0]) != 7) fail(GNSSACQ_ERR_INVALID_ARG, "channels must be a 7 x n double matrix");
        ch = (gnssacq_channel*)mxMalloc(sizeof(gnssacq_channel) * (size_t)(n_ch + 1));
        oi = (double*)mxMalloc(sizeof(double) * (size_t)(n_ch * n_taps + 1));
        oq = (double*)mxMalloc(sizeof(double) * (size_t)(n_ch * n_taps + 1));
        for (i = 0; i < n_ch; ++i) {
            ch[i].prn = (int32_t)m[7 * i];
            ch[i].num_samples = (int32_t)m[7 * i + 1];
            ch[i].sample_offset = (int64_t)m[7 * i + 2];
            ch[i].carrier_hz = m[7 * i + 3];
            ch[i].rem_phase = m[7 * i + 4];
            ch[i].code_hz = m[7 * i + 5];
            ch[i].rem_chip = m[7 * i + 6];
        }
        rc = gnssacq_correlate(g_handle, n_ch, ch, n_taps, mxGetPr(prhs[2]), oi, oq);
        if (rc != GNSSACQ_OK) { mxFree(ch); mxFree(oi); mxFree(oq); fail(rc, gnssacq_last_error(g_handle)); }
        plhs[0] = mxCreateDoubleMatrix((mwSize)n_ch, (mwSize)n_taps, mxREAL);
        pi = mxGetPr(plhs[0]);
        if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix((mwSize)n_ch, (mwSize)n_taps, mxREAL); pq = mxGetPr(plhs[1]); } else pq = NULL;
        for (i = 0; i < n_ch; ++i)
            for (t = 0; t < n_taps; ++t) {             /* row-major [channel][tap] -> MATLAB column-major */
                pi[i + t * n_ch] = oi[i * n_taps + t];
                if (pq) pq[i + t * n_ch] = oq[i * n_taps + t];
            }
        mxFree(ch); mxFree(oi); mxFree(oq);
        return;
    }
    if (track_mode) {                                  /* whole conventional loop (trackingCT.m:70-172) */
        const double* m = mxGetPr(prhs[0]);
        const double* lp = mxGetPr(prhs[2]);
        const int n_ch = (int)mxGetN(prhs[0]), n_per = (int)mxGetScalar(prhs[3]);
        gnssacq_channel* ch;
        gnssacq_track_record* rec;
        gnssacq_loop_params loops;
        mwSize dims[3];
        int k;
        if (!mxIsDouble(prhs[0]) || mxGetM(prhs[0]) != 7 || mxGetNumberOfElements(prhs[2]) != 7 || n_per < 1)
            fail(GNSSACQ_ERR_INVALID_ARG, "track: channels 7 x n, loops 1 x 7, n_periods >= 1");
        loops.dll_bw = lp[0]; loops.dll_damp = lp[1]; loops.dll_gain = lp[2];
        loops.pll_bw = lp[3]; loops.pll_damp = lp[4]; loops.pll_gain = lp[5]; loops.spacing_chips = lp[6];
        ch = (gnssacq_channel*)mxMalloc(sizeof(gnssacq_channel) * (size_t)(n_ch + 1));
        rec = (gnssacq_track_record*)mxMalloc(sizeof(gnssacq_track_record) * ((size_t)n_ch * (size_t)n_per + 1));
        for (i = 0; i < n_ch; ++i) {
            ch[i].prn = (int32_t)m[7 * i];
            ch[i].num_samples = 0;
            ch[i].sample_offset = (int64_t)m[7 * i + 2];
            ch[i].carrier_hz = m[7 * i + 3];
            ch[i].rem_phase = m[7 * i + 4];
            ch[i].code_hz = m[7 * i + 5];
            ch[i].rem_chip = m[7 * i + 6];
        }
        rc = gnssacq_track(g_handle, n_ch, ch, &loops, n_per, rec);
        if (rc != GNSSACQ_OK) { mxFree(ch); mxFree(rec); fail(rc, gnssacq_last_error(g_handle)); }
        dims[0] = 14; dims[1] = (mwSize)n_per; dims[2] = (mwSize)n_ch;
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        out = mxGetPr(plhs[0]);
        for (i = 0; i < n_ch; ++i)
            for (k = 0; k < n_per; ++k) {
                const gnssacq_track_record* r = &rec[(size_t)i * n_per + k];
                double* o = out + 14 * ((size_t)i * n_per + k);
                o[0] = r->P_i; o[1] = r->P_q; o[2] = r->E_i; o[3] = r->E_q; o[4] = r->L_i; o[5] = r->L_q;
                o[6] = r->pll_discri; o[7] = r->dll_discri; o[8] = r->rem_chip; o[9] = r->code_hz;
                o[10] = r->carrier_hz; o[11] = r->rem_phase; o[12] = (double)r->sample_end; o[13] = r->num_samples;
            }
        mxFree(ch); mxFree(rec);
        return;
    }
    if (nrhs == 5) {                                   /* fine-frequency stage (acquisition.m:83-127) */
        int n_sv = (int)mxGetNumberOfElements(prhs[3]);
        const double* svd = mxGetPr(prhs[3]);
        const double* cdd = mxGetPr(prhs[4]);
        int32_t* sv = (int32_t*)mxMalloc(sizeof(int32_t) * (size_t)(n_sv + 1));
        int32_t* cd = (int32_t*)mxMalloc(sizeof(int32_t) * (size_t)(n_sv + 1));
        if ((int)mxGetNumberOfElements(prhs[4]) != n_sv) fail(GNSSACQ_ERR_INVALID_ARG, "sv and codedelay differ in length");
        for (i = 0; i < n_sv; ++i) { sv[i] = (int32_t)svd[i]; cd[i] = (int32_t)cdd[i]; }
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)n_sv, mxREAL);
        rc = gnssacq_fine_frequency(g_handle, mxGetData(prhs[0]), nbytes, (int32_t)mxGetScalar(prhs[2]), n_sv,
                                    sv, cd, mxGetPr(plhs[0]));
        mxFree(sv);
        mxFree(cd);
        if (rc != GNSSACQ_OK) fail(rc, gnssacq_last_error(g_handle));
        return;
    }
    rows = (gnssacq_result*)mxMalloc(sizeof(gnssacq_result) * (size_t)c.n_prn);
    rc = gnssacq_search(g_handle, mxGetData(prhs[0]), nbytes, rows, NULL);
    if (rc != GNSSACQ_OK) { mxFree(rows); fail(rc, gnssacq_last_error(g_handle)); }

    plhs[0] = mxCreateDoubleMatrix((mwSize)c.n_prn, 8, mxREAL);
    out = mxGetPr(plhs[0]);
    for (i = 0; i < c.n_prn; ++i) {
        out[i + 0 * c.n_prn] = rows[i].prn;
        out[i + 1 * c.n_prn] = rows[i].acquired;
        out[i + 2 * c.n_prn] = rows[i].code_phase;
        out[i + 3 * c.n_prn] = rows[i].doppler_bin;
        out[i + 4 * c.n_prn] = rows[i].doppler_hz;
        out[i + 5 * c.n_prn] = rows[i].peak;
        out[i + 6 * c.n_prn] = rows[i].noise_meansq;
        out[i + 7 * c.n_prn] = rows[i].snr_db;
    }
    mxFree(rows);
    (void)nlhs;
}
