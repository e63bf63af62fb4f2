// gnss_cplx.h -- complex helpers, compile-time loops and compile-time twiddles.
//
// Everything here is `GNSS_HD`: it compiles as __host__ __device__ under nvcc
// and as plain inline C++17 under g++.  The g++ build exists only so the
// engine's index arithmetic can be exercised by the CPU test-suite
// (tests/emu/); the product path is the CUDA build.
#pragma once
#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
#define GNSS_HD __host__ __device__ __forceinline__
#else
#define GNSS_HD inline
#endif

namespace gnss {

struct alignas(8) cf {
    float x, y;
};

GNSS_HD cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }

// Complex arithmetic.  On the device every helper is ONE packed FP32 instruction of sm_100a
// (FADD2 / FMUL2 / FFMA2 through __fadd2_rn / __fmul2_rn / __ffma2_rn): a complex value is a register
// pair, the instruction's operand modifiers provide negation, the re<->im swap (.LO_HI) and the
// broadcast of a scalar register or of a 32-bit immediate.  Same FMA-pipe time as two scalar
// instructions, half the issue slots and half the code bytes (profiles/r02/fp32x2_probe.txt).
// The host forms (CPU emulation of the index maps, tests/emu) compute the same values unfused.
#if defined(__CUDA_ARCH__)
#define GNSS_PACKED 1
__device__ __forceinline__ float2 f2(cf a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cf fc(float2 a) { return mk(a.x, a.y); }
#endif

GNSS_HD cf cadd(cf a, cf b) {
#ifdef GNSS_PACKED
    return fc(__fadd2_rn(f2(a), f2(b)));
#else
    return mk(a.x + b.x, a.y + b.y);
#endif
}
GNSS_HD cf csub(cf a, cf b) {
#ifdef GNSS_PACKED
    return fc(__fadd2_rn(f2(a), make_float2(-b.x, -b.y)));
#else
    return mk(a.x - b.x, a.y - b.y);
#endif
}
// a * c, c real
GNSS_HD cf cscale(cf a, float c) {
#ifdef GNSS_PACKED
    return fc(__fmul2_rn(f2(a), make_float2(c, c)));
#else
    return mk(a.x * c, a.y * c);
#endif
}
// acc + a * c, c real
GNSS_HD cf cfma(cf a, float c, cf acc) {
#ifdef GNSS_PACKED
    return fc(__ffma2_rn(f2(a), make_float2(c, c), f2(acc)));
#else
    return mk(acc.x + a.x * c, acc.y + a.y * c);
#endif
}
// a - i*b  and  a + i*b
GNSS_HD cf caddmi(cf a, cf b) {
#ifdef GNSS_PACKED
    return fc(__ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), f2(a)));
#else
    return mk(a.x + b.y, a.y - b.x);
#endif
}
GNSS_HD cf caddpi(cf a, cf b) {
#ifdef GNSS_PACKED
    return fc(__ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), f2(a)));
#else
    return mk(a.x - b.y, a.y + b.x);
#endif
}
GNSS_HD cf cmul(cf a, cf b) {
#ifdef GNSS_PACKED
    const float2 t = __fmul2_rn(make_float2(a.x, a.x), f2(b));
    return fc(__ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), t));
#else
    return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
// Scalar form (2 FMUL + 2 FFMA = 4 FMA-pipe cycles; the packed form above is 3 instructions but 5 pipe
// cycles): used where the FMA pipe, not the issue slot, is the limiter (pass 1 of the search).
GNSS_HD cf cmul_scalar(cf a, cf b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// v * (c - i s) with (c, s) known: c*v + s*(v.y, -v.x)
GNSS_HD cf cmul_cs(cf v, float c, float s) {
#ifdef GNSS_PACKED
    const float2 t = __fmul2_rn(f2(v), make_float2(c, c));
    return fc(__ffma2_rn(make_float2(v.y, v.x), make_float2(s, -s), t));
#else
    return mk(v.x * c + v.y * s, v.y * c - v.x * s);
#endif
}
// acc + |a|^2.  |a|^2 is rounded on its own before the addition: a block's power plane published by the
// block-granular tail (search_kernel_coop) and added later must give the bits of the in-place accumulation.
GNSS_HD float cnorm_acc(cf a, float acc) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(acc, fmaf(a.y, a.y, a.x * a.x));
#else
    return acc + (a.x * a.x + a.y * a.y);
#endif
}
GNSS_HD float cnorm(cf a) { return a.x * a.x + a.y * a.y; }
// read-only global load of one complex value (LDG.E.64.CONSTANT on the device)
GNSS_HD cf ld_ro(const cf* p) {
#if defined(__CUDA_ARCH__) && defined(GNSS_EXPERIMENT_LDNA)     // experiment: operands do not allocate in L1 (each line is used once)
    float2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return mk(v.x, v.y);
#elif defined(__CUDA_ARCH__)
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return mk(v.x, v.y);
#else
    return *p;
#endif
}

// 16-byte pair of complex values; ld_cg2 = L2-only (cache-global) load, used for data another SM wrote
struct alignas(16) cf2 {
    cf lo, hi;
};
GNSS_HD cf2 ld_cg2(const cf2* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    cf2 r;
    r.lo = mk(v.x, v.y);
    r.hi = mk(v.z, v.w);
    return r;
#else
    return *p;
#endif
}

// ---- compile-time loop: f(std::integral_constant<int, I>) for I in [B, E) ----
template <int B, int E, class F>
GNSS_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- compile-time trigonometry (double Taylor series on a reduced argument) ----
namespace detail {
constexpr double kPi = 3.14159265358979323846264338327950288;
// sin/cos of 2*pi*num/den, |result| exact to ~1e-16; octant symmetries are
// resolved in integer arithmetic so special angles come out exact.
constexpr double taylor_sin(double x) {   // |x| <= pi/4
    double term = x, sum = x, x2 = x * x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double taylor_cos(double x) {
    double term = 1.0, sum = 1.0, x2 = x * x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
struct sc { double s, c; };
constexpr sc sincos_turn(long long num, long long den) {   // angle = 2*pi*num/den
    num %= den; if (num < 0) num += den;
    // work in eighths of a turn: t = num/den in [0,1); octant o = floor(8t)
    long long o = (8 * num) / den;
    long long rnum = 8 * num - o * den;            // remainder/ (8 den) of a turn, in [0, den)
    // angle inside the octant: phi = 2*pi * rnum / (8 den) in [0, pi/4)
    double phi = 2.0 * kPi * (double)rnum / (8.0 * (double)den);
    double s = taylor_sin(phi), c = taylor_cos(phi);
    if (rnum == 0) { s = 0.0; c = 1.0; }
    // rotate by o * 45 degrees
    constexpr double h = 0.70710678118654752440084436210484903928;
    double s1 = 0, c1 = 0;
    switch (o) {
        case 0: s1 = s; c1 = c; break;
        case 1: s1 = h * (s + c); c1 = h * (c - s); break;
        case 2: s1 = c; c1 = -s; break;
        case 3: s1 = h * (c - s); c1 = -h * (c + s); break;
        case 4: s1 = -s; c1 = -c; break;
        case 5: s1 = -h * (s + c); c1 = h * (s - c); break;
        case 6: s1 = -c; c1 = s; break;
        default: s1 = h * (s - c); c1 = h * (c + s); break;
    }
    return sc{s1, c1};
}
}  // namespace detail

// cos / sin of 2*pi*NUM/DEN as float compile-time constants.
template <int NUM, int DEN>
struct Tw {
    static constexpr float c = (float)detail::sincos_turn(NUM, DEN).c;
    static constexpr float s = (float)detail::sincos_turn(NUM, DEN).s;
};

// v * exp(-2*pi*i*NUM/DEN)  (forward-transform twiddle), special-casing 1/8 turns.
template <int NUM, int DEN>
GNSS_HD cf mul_tw(cf v) {
    constexpr int T = ((NUM % DEN) + DEN) % DEN;
    constexpr float h = 0.70710678118654752440f;
    if constexpr (T == 0) {
        return v;
    } else if constexpr (4 * T == DEN) {          // -i
        return caddmi(mk(0.f, 0.f), v);
    } else if constexpr (2 * T == DEN) {          // -1
        return mk(-v.x, -v.y);
    } else if constexpr (4 * T == 3 * DEN) {      // +i
        return caddpi(mk(0.f, 0.f), v);
    } else if constexpr (8 * T == DEN) {          // (1-i)/sqrt2 : h * (v - i v)
        return cscale(caddmi(v, v), h);
    } else if constexpr (8 * T == 3 * DEN) {      // (-1-i)/sqrt2 : -h * (v + i v)
        return cscale(caddpi(v, v), -h);
    } else {
        return cmul_cs(v, Tw<T, DEN>::c, Tw<T, DEN>::s);
    }
}

// ---- small compile-time integer helpers ----
constexpr int cmodinv(int a, int m) {
    a %= m;
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 0;
}

}  // namespace gnss
