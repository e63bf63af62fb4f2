// CPU emulation of the CUDA engine -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the very same `GNSS_HD` pass functions the kernels call
// (csrc/gnss_engine.h) with g++ and runs them task by task, cluster CTA by
// cluster CTA, so the prime-factor index maps, layouts, bin-shift identity and
// functors can be checked against NumPy without a GPU.  It is never loaded by
// the product (libgnssacq.so has no CPU path).
#include <cstring>
#include <vector>
#include "gnss_engine.h"

using namespace gnss;

template <int Q, int R, class Loader, class Storer>
static void run_unit(const Loader& ld, Storer* st /*R storers*/, std::vector<std::vector<cf>>& D,
                     const std::vector<Tw4>& tw) {
    using S = Split<Q, R>;
    cf* Dall[R];
    for (int r = 0; r < R; ++r) Dall[r] = D[r].data();
    for (int r = 0; r < R; ++r)
        for (int t = 0; t < S::P1_TASKS; ++t) pass1_task<Q, R>(t, r, ld, Dall[r]);
    for (int r = 0; r < R; ++r)                      // the CTA-wide form the kernels call, thread by thread
        for (int tid = 0; tid < 128; ++tid) pass2_cta<Q, R, 128>(tid, Dall[r], tw.data());
    for (int r = 0; r < R; ++r)
        for (int t = 0; t < S::P3_TASKS; ++t) pass3_task<Q, R>(t, Dall[r]);
    for (int r = 0; r < R; ++r)
        for (int t = 0; t < S::P4_TASKS; ++t) pass4_task<Q, R>(t, r, Dall, st[r]);
}

static std::vector<Tw4> make_tw125() {
    std::vector<Tw4> tw(100);                        // entry b2*4 + (k1-1): W125^(b2*k1), as fill_tw125 builds it
    for (int j = 0; j < 100; ++j) {
        double a = -2.0 * 3.14159265358979323846 * ((j >> 2) * ((j & 3) + 1)) / 125.0;
        tw[j] = Tw4{(float)cos(a), 0.f, -(float)sin(a), (float)sin(a)};
    }
    return tw;
}

template <int Q, int R>
static int code_spectrum(const int8_t* scode, cf* out) {
    using S = Split<Q, R>;
    std::vector<std::vector<cf>> D(R, std::vector<cf>(S::D_ELEMS));
    auto tw = make_tw125();
    CodeLoader ld{scode};
    SpectrumStorer st[R];
    for (int r = 0; r < R; ++r) st[r] = SpectrumStorer{out, 1.0f / Geo<Q>::N, 1, 0};
    run_unit<Q, R>(ld, st, D, tw);
    return 0;
}

template <int Q, int R>
static int wipe_spectrum(const void* raw, int data_type, int precision, int coh_ms, double f_hz,
                         double fs_hz, float mi, float mq, cf* out) {
    using S = Split<Q, R>;
    std::vector<std::vector<cf>> D(R, std::vector<cf>(S::D_ELEMS));
    auto tw = make_tw125();
    WipeoffLoader ld{raw, data_type, precision, coh_ms, f_hz, fs_hz, mi, mq};
    SpectrumStorer st[R];
    for (int r = 0; r < R; ++r) st[r] = SpectrumStorer{out, 1.0f, 0, 1};
    run_unit<Q, R>(ld, st, D, tw);
    return 0;
}

template <int Q, int R>
static int search_row(const cf* cc, const cf* x_blocks, int K, int shift, float* acc_by_lag) {
    using S = Split<Q, R>;
    using G = Geo<Q>;
    std::vector<std::vector<cf>> D(R, std::vector<cf>(S::D_ELEMS));
    std::vector<std::vector<float>> acc(R, std::vector<float>(S::ACC_ELEMS, 0.f));
    auto tw = make_tw125();
    int sa, sb, sc;
    G::shift_coords(shift, sa, sb, sc);
    for (int k = 0; k < K; ++k) {
        SearchLoader ld{cc, x_blocks + (size_t)k * G::NX, sa, sb, sc};
        PowerAccumStorer st[R];
        for (int r = 0; r < R; ++r) st[r] = PowerAccumStorer{acc[r].data()};
        run_unit<Q, R>(ld, st, D, tw);
    }
    for (int r = 0; r < R; ++r)
        for (int ap = 0; ap < 16; ++ap)
            for (int t = 0; t < S::CH; ++t) {
                int col = r * S::CH + t;
                if (col >= S::ROW) continue;
                acc_by_lag[G::lag_of(ap, col)] = acc[r][ap * S::CH + t];
            }
    // self-check of cell_of_lag (inverse of lag_of)
    for (int m = 0; m < G::N; ++m) {
        int ap, col;
        G::cell_of_lag(m, ap, col);
        if (G::lag_of(ap, col) != m) return -1;
    }
    return 0;
}

template <int Q, int R>
static int fine_unit(const void* raw, const uint16_t* chip, const int8_t* ca, int data_type, int precision,
                     float mi, float mq, int start, int L, int n2, int r, long long F, cf* out) {
    using S = Split<Q, R>;
    std::vector<std::vector<cf>> D(R, std::vector<cf>(S::D_ELEMS));
    auto tw = make_tw125();
    FineLoader ld;
    ld.raw = raw; ld.chip = chip; ld.ca = ca; ld.data_type = data_type; ld.precision = precision;
    ld.mean_i = mi; ld.mean_q = mq; ld.start = start; ld.L = L; ld.n2 = n2; ld.r = r; ld.F = F;
    NaturalStorerHD st[R];
    for (int k = 0; k < R; ++k) st[k] = NaturalStorerHD{out};
    run_unit<Q, R>(ld, st, D, tw);
    return 0;
}

template <int Q>
static void gx_to_natural(const cf* g, cf* nat) {   // c-extended layout [16][2Q-1][125]; also checks the duplicate planes
    using G = Geo<Q>;
    for (int a = 0; a < 16; ++a)
        for (int b = 0; b < 125; ++b)
            for (int c = 0; c < Q; ++c) {
                cf v = g[(a * (2 * Q - 1) + c) * 125 + b];
                if (c + Q < 2 * Q - 1) {
                    cf w = g[(a * (2 * Q - 1) + c + Q) * 125 + b];
                    if (w.x != v.x || w.y != v.y) v = mk(1e30f, 1e30f);
                }
                nat[G::good(a, b, c)] = v;
            }
}
// search row through the transposed (XT) exchange path of the cooperative kernel
template <int Q, int R>
static int search_row_xt(const cf* cc, const cf* x_blocks, int K, int shift, float* acc_by_lag) {
    using S = Split<Q, R>;
    using G = Geo<Q>;
    using GX = GeoX<Q>;
    using SX = SplitX<Q, R>;
    std::vector<std::vector<cf>> D(R, std::vector<cf>(S::D_ELEMS));
    std::vector<std::vector<float>> acc(R, std::vector<float>(SX::ACC_ELEMS, 0.f));
    std::vector<cf> xbuf((size_t)16 * GX::RSX, mk(1e30f, 1e30f));
    auto tw = make_tw125();
    int sa, sb, sc;
    G::shift_coords(shift, sa, sb, sc);
    for (int k = 0; k < K; ++k) {
        SearchLoader ld{cc, x_blocks + (size_t)k * G::NX, sa, sb, sc};
        for (int r = 0; r < R; ++r) {
            for (int t = 0; t < S::P1_TASKS; ++t) pass1_task<Q, R>(t, r, ld, D[r].data());
            for (int t = 0; t < S::P2_TASKS; ++t) pass2_task<Q, R>(t, D[r].data(), tw.data());
            for (int t = 0; t < S::P3_TASKS; ++t) pass3_task_xt<Q, R>(t, D[r].data(), xbuf.data() + (size_t)r * S::A * GX::RSX);
        }
        for (int r = 0; r < R; ++r) {
            PowerAccumStorer st{acc[r].data()};
            for (int j = 0; j < SX::P4_TASKS; ++j) pass4_task_xt<Q, R>(j, r, xbuf.data(), st);
        }
    }
    for (int m = 0; m < G::N; ++m) acc_by_lag[m] = -1.f;
    for (int r = 0; r < R; ++r)
        for (int ap = 0; ap < 16; ++ap)
            for (int t = 0; t < SX::CHX; ++t) {
                const int e = r * SX::CHX + t;
                if (!GX::valid(e)) continue;
                acc_by_lag[GX::lag_of(ap, e)] = acc[r][ap * SX::CHX + t];
            }
    for (int m = 0; m < G::N; ++m) {
        int ap, e;
        GX::cell_of_lag(m, ap, e);
        if (!GX::valid(e) || GX::lag_of(ap, e) != m) return -1;
    }
    return 0;
}

template <int Q>
static void g_to_natural(const cf* g, cf* nat) {
    using G = Geo<Q>;
    for (int a = 0; a < 16; ++a)
        for (int b = 0; b < 125; ++b)
            for (int c = 0; c < Q; ++c) nat[G::good(a, b, c)] = g[G::gidx(a, b, c)];
}

#define DISPATCH(Qv, Rv, CALL)                                         \
    if (Q == Qv && R == Rv) { constexpr int QQ = Qv, RR = Rv; (void)QQ; (void)RR; return CALL; }
#define ALL(CALLQ)                                                     \
    DISPATCH(3, 1, CALLQ) DISPATCH(3, 2, CALLQ) DISPATCH(3, 4, CALLQ)  \
    DISPATCH(13, 2, CALLQ) DISPATCH(13, 4, CALLQ) DISPATCH(29, 4, CALLQ) DISPATCH(29, 8, CALLQ)

extern "C" {
int emu_code_spectrum(int Q, int R, const int8_t* scode, float* out) {
    ALL((code_spectrum<QQ, RR>(scode, (cf*)out)))
    return -2;
}
int emu_wipe_spectrum(int Q, int R, const void* raw, int data_type, int precision, int coh_ms,
                      double f_hz, double fs_hz, float mi, float mq, float* out) {
    ALL((wipe_spectrum<QQ, RR>(raw, data_type, precision, coh_ms, f_hz, fs_hz, mi, mq, (cf*)out)))
    return -2;
}
int emu_search_row(int Q, int R, const float* cc, const float* x_blocks, int K, int shift,
                   float* acc_by_lag) {
    ALL((search_row<QQ, RR>((const cf*)cc, (const cf*)x_blocks, K, shift, acc_by_lag)))
    return -2;
}
int emu_fine_unit(int Q, int R, const void* raw, const uint16_t* chip, const int8_t* ca, int data_type,
                  int precision, float mi, float mq, int start, int L, int n2, int r, long long F, float* out) {
    ALL((fine_unit<QQ, RR>(raw, chip, ca, data_type, precision, mi, mq, start, L, n2, r, F, (cf*)out)))
    return -2;
}
int emu_gx_to_natural(int Q, const float* g, float* nat) {
    if (Q == 3) { gx_to_natural<3>((const cf*)g, (cf*)nat); return 0; }
    if (Q == 13) { gx_to_natural<13>((const cf*)g, (cf*)nat); return 0; }
    if (Q == 29) { gx_to_natural<29>((const cf*)g, (cf*)nat); return 0; }
    return -2;
}
int emu_search_row_xt(int Q, int R, const float* cc, const float* x_blocks, int K, int shift, float* acc_by_lag) {
    ALL((search_row_xt<QQ, RR>((const cf*)cc, (const cf*)x_blocks, K, shift, acc_by_lag)))
    return -2;
}
int emu_g_to_natural(int Q, const float* g, float* nat) {
    if (Q == 3) { g_to_natural<3>((const cf*)g, (cf*)nat); return 0; }
    if (Q == 13) { g_to_natural<13>((const cf*)g, (cf*)nat); return 0; }
    if (Q == 29) { g_to_natural<29>((const cf*)g, (cf*)nat); return 0; }
    return -2;
}
}
