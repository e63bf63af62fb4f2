// gnss_radix.h -- register-resident forward DFT butterflies (exp(-2*pi*i*nk/R)).
//
//   dft_odd<Q>   any odd length (5, 13, 29, ...): symmetric-pair form, all coefficients compile-time
//                immediates of packed FFMA2 instructions: 3H + 2H*H + 2H packed ops, H=(Q-1)/2
//   dft4, dft16  16 = 4 x 4 Cooley-Tukey
//   dft25        25 = 5 x 5 Cooley-Tukey
//
// dft16 / dft25 leave their result digit-reversed: v[R1*k1 + k2] = X[k1 + R1*k2]
// (R1 = 4 resp. 5); callers index accordingly, which costs nothing because all
// register indices are static.
#pragma once
#include "gnss_cplx.h"

namespace gnss {

template <int Q>
GNSS_HD void dft_odd(cf (&v)[Q]) {
    static_assert(Q % 2 == 1 && Q >= 3, "odd length");
    constexpr int H = (Q - 1) / 2;
    cf a[H + 1], b[H + 1];
    const cf x0 = v[0];
    cf s0 = x0;
    static_for<1, H + 1>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        a[J] = cadd(v[J], v[Q - J]);
        b[J] = csub(v[J], v[Q - J]);
        s0 = cadd(s0, a[J]);
    });
    v[0] = s0;
    static_for<1, H + 1>([&](auto kc) {
        constexpr int K = decltype(kc)::value;
        // C = x0 + sum_J cos(2 pi JK/Q) a[J],  S = sum_J sin(2 pi JK/Q) b[J]  (one packed FMA per term)
        cf C = x0, S = mk(0.f, 0.f);
        static_for<1, H + 1>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            constexpr int T = (J * K) % Q;
            C = cfma(a[J], Tw<T, Q>::c, C);
            if constexpr (J == 1) S = cscale(b[J], Tw<T, Q>::s);
            else S = cfma(b[J], Tw<T, Q>::s, S);
        });
        // X[K] = C - i S ; X[Q-K] = C + i S
        v[K] = caddmi(C, S);
        v[Q - K] = caddpi(C, S);
    });
}

// Same transform, input-major ("streaming") form: inputs are fetched pair by pair through `get`
// (get(integral_constant<int,C>) -> cf) and folded into all (Q-1)/2 output accumulators at once, so the
// arithmetic on the first pairs overlaps the memory latency of the later ones.  2*H complex accumulators live.
template <int Q, class Get>
GNSS_HD void dft_odd_stream(Get&& get, cf (&v)[Q]) {
    static_assert(Q % 2 == 1 && Q >= 3, "odd length");
    constexpr int H = (Q - 1) / 2;
    const cf x0 = get(std::integral_constant<int, 0>{});
    cf C[H + 1], S[H + 1];
    cf s0 = x0;
    static_for<1, H + 1>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        const cf p = get(std::integral_constant<int, J>{});
        const cf q = get(std::integral_constant<int, Q - J>{});
        const cf a = cadd(p, q), b = csub(p, q);
        s0 = cadd(s0, a);
        static_for<1, H + 1>([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int T = (J * K) % Q;
            if constexpr (J == 1) {
                C[K] = cfma(a, Tw<T, Q>::c, x0);
                S[K] = cscale(b, Tw<T, Q>::s);
            } else {
                C[K] = cfma(a, Tw<T, Q>::c, C[K]);
                S[K] = cfma(b, Tw<T, Q>::s, S[K]);
            }
        });
    });
    v[0] = s0;
    static_for<1, H + 1>([&](auto kc) {
        constexpr int K = decltype(kc)::value;
        v[K] = caddmi(C[K], S[K]);
        v[Q - K] = caddpi(C[K], S[K]);
    });
}

GNSS_HD void dft4(cf& x0, cf& x1, cf& x2, cf& x3) {
    const cf t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = caddmi(t1, t3);   // t1 - i t3
    x3 = caddpi(t1, t3);   // t1 + i t3
}
// dft4 whose input x2 still has to be multiplied by -i (the W16^4 twiddle): folded into the first butterfly
GNSS_HD void dft4_x2mi(cf& x0, cf& x1, cf& x2, cf& x3) {
    const cf t0 = caddmi(x0, x2), t1 = caddpi(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = caddmi(t1, t3);
    x3 = caddpi(t1, t3);
}

// in: v[4*n1 + n2] = x[4*n1 + n2]; out: v[4*k1 + k2] = X[k1 + 4*k2]
GNSS_HD void dft16(cf (&v)[16]) {
    static_for<0, 4>([&](auto n2c) {
        constexpr int N2 = decltype(n2c)::value;
        dft4(v[N2], v[4 + N2], v[8 + N2], v[12 + N2]);     // over n1 -> k1 at v[4*k1 + N2]
        static_for<1, 4>([&](auto k1c) {
            constexpr int K1 = decltype(k1c)::value;
            if constexpr (N2 * K1 != 4) v[4 * K1 + N2] = mul_tw<N2 * K1, 16>(v[4 * K1 + N2]);
        });
    });
    static_for<0, 4>([&](auto k1c) {
        constexpr int K1 = decltype(k1c)::value;
        // over n2 -> k2; v[4*2 + 2] carries the pending -i (N2*K1 == 4)
        if constexpr (K1 == 2) dft4_x2mi(v[8], v[9], v[10], v[11]);
        else dft4(v[4 * K1], v[4 * K1 + 1], v[4 * K1 + 2], v[4 * K1 + 3]);
    });
}

// in: v[5*n1 + n2]; out: v[5*k1 + k2] = X[k1 + 5*k2]
GNSS_HD void dft25(cf (&v)[25]) {
    static_for<0, 5>([&](auto n2c) {
        constexpr int N2 = decltype(n2c)::value;
        cf u[5] = {v[N2], v[5 + N2], v[10 + N2], v[15 + N2], v[20 + N2]};
        dft_odd<5>(u);
        v[N2] = u[0];
        static_for<1, 5>([&](auto k1c) {
            constexpr int K1 = decltype(k1c)::value;
            v[5 * K1 + N2] = mul_tw<N2 * K1, 25>(u[K1]);
        });
    });
    static_for<0, 5>([&](auto k1c) {
        constexpr int K1 = decltype(k1c)::value;
        cf u[5] = {v[5 * K1], v[5 * K1 + 1], v[5 * K1 + 2], v[5 * K1 + 3], v[5 * K1 + 4]};
        dft_odd<5>(u);
        static_for<0, 5>([&](auto k2c) {
            constexpr int K2 = decltype(k2c)::value;
            v[5 * K1 + K2] = u[K2];
        });
    });
}

}  // namespace gnss
