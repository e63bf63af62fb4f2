// tc5_dft29_probe.cu -- the radix-29 stage of pass 1 as a GEMM on the 5th-generation tensor cores (tcgen05.mma,
// kind::tf32, accumulator in tensor memory), against the product's packed-FP32 butterflies (gnss::dft_odd<29>).
//
// BASELINE.json north_star: "Tensor cores are used only if a DFT-as-GEMM radix stage is shown by ncu to beat the
// CUDA-core butterflies."  r01 measured the legacy mma.sync path (profiles/tc_dft29_probe.cu: 1.46x slower).  This is
// the sm_100a path proper:
//   * one CTA tile = 128 transform instances ("columns" of the prime-factor array) = the M = 128 rows of the MMA;
//   * A[128 x 64]: row i = (re z_0..z_28, im z_0..z_28, 6 zeros) of instance i, TF32, K-major, no swizzle;
//   * B[64 x 64]:  the real form of W29 (cos / sin / -sin / cos blocks), K-major;  D[128 x 64] FP32 in TMEM;
//   * FP32 accuracy through the error-compensated 3xTF32 split: D = Ahi*Bhi + Alo*Bhi + Ahi*Blo
//     (24 tcgen05.mma of M128 N64 K8 per tile, issued by ONE thread; tcgen05.commit -> mbarrier);
//   * operands are written to shared memory by the threads that own the rows (in the kernel they are products
//     cc*x formed in registers, not something a TMA could fetch), results come back with tcgen05.ld.
// Reports columns/s of both paths on register-resident inputs with the product's shared-memory stores, and the
// accuracy of both against a float64 DFT.  Not product code.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../assignment-for-aae6102_gnss-sdr_b200/csrc tc5_dft29_probe.cu -o tc5_dft29_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "gnss_radix.h"

using gnss::cf;
constexpr int Q = 29, TM = 128, KD = 64, ND = 64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- A: CUDA-core butterflies (packed FP32)
__global__ void __launch_bounds__(128, 3) dft29_ffma(const cf* __restrict__ x, cf* __restrict__ y, int iters, float s0) {
    extern __shared__ __align__(16) unsigned char dyn[];
    cf* in = reinterpret_cast<cf*>(dyn);
    cf* out = in + 128 * 29;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < Q; ++c) in[c * 128 + threadIdx.x] = x[(size_t)col * Q + c];
    float s = s0;
    for (int it = 0; it < iters; ++it) {
        cf v[Q];
#pragma unroll
        for (int c = 0; c < Q; ++c) { const cf t = in[c * 128 + threadIdx.x]; v[c] = gnss::mk(t.x * s, t.y * s); }
        gnss::dft_odd<Q>(v);
#pragma unroll
        for (int c = 0; c < Q; ++c) out[c * 128 + threadIdx.x] = v[c];
        // the next iteration's scale depends on this iteration's outputs: nothing can be hoisted or dropped
        s += 1.0f + (v[1].x + v[7].y + v[28].x) * 1e-30f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < Q; ++c) y[(size_t)col * Q + c] = out[c * 128 + threadIdx.x];
}

// ---------------------------------------------------------------- B: tcgen05.mma, 3xTF32
__device__ __forceinline__ float tf32_rn(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes, stored contiguously
// (128 B); lbo = byte distance between core matrices adjacent in K, sbo = between 8-row groups (M / N direction)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version of sm_100
    return d;                                     // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor of kind::tf32: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ND >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}

// smem: [A_hi 32 KB][A_lo 32 KB][B_hi 16 KB][B_lo 16 KB]; A: element (i, k) at (k/4)*2048 + (i/8)*128 + (i%8)*16 + (k%4)*4
// (lbo 2048, sbo 128); B: element (n, k) at (k/4)*1024 + (n/8)*128 + (n%8)*16 + (k%4)*4 (lbo 1024, sbo 128)
__global__ void __launch_bounds__(128, 2) dft29_tc5(const cf* __restrict__ x, cf* __restrict__ y, const float* __restrict__ bmat /*[2][64 n][64 k]*/,
                                                    int iters, float s0) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_slot;
    unsigned char* a_hi = dyn;
    unsigned char* a_lo = dyn + 32768;
    unsigned char* b_hi = dyn + 65536;
    unsigned char* b_lo = dyn + 65536 + 16384;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int col = blockIdx.x * blockDim.x + tid;
    // B matrices once
    for (int e = tid; e < 2 * ND * KD; e += 128) {
        const int m = e / (ND * KD), r = e - m * ND * KD, n = r / KD, k = r - n * KD;
        unsigned char* dst = (m ? b_lo : b_hi) + (k >> 2) * 1024 + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(dst) = bmat[e];
    }
    if (tid == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(&mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(&tmem_slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "n"(ND) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    const uint32_t mine = tbase + ((uint32_t)(32 * warp) << 16);
    cf z0[Q];
#pragma unroll
    for (int c = 0; c < Q; ++c) z0[c] = x[(size_t)col * Q + c];
    const uint32_t sa_hi = (uint32_t)__cvta_generic_to_shared(a_hi), sa_lo = (uint32_t)__cvta_generic_to_shared(a_lo);
    const uint32_t sb_hi = (uint32_t)__cvta_generic_to_shared(b_hi), sb_lo = (uint32_t)__cvta_generic_to_shared(b_lo);
    const uint32_t smbar = (uint32_t)__cvta_generic_to_shared(&mbar);
    unsigned char* row_hi = a_hi + (tid >> 3) * 128 + (tid & 7) * 16;
    unsigned char* row_lo = a_lo + (tid >> 3) * 128 + (tid & 7) * 16;
    // K = 64 per row: k = c (re), 29 + c (im), 58..63 zero -- 16 chunks of 4
    float s = s0;
    uint32_t phase = 0;
    float acc_chk = 0.f;
    float res[ND];
    for (int it = 0; it < iters; ++it) {
        // ---- operand staging: hi / lo split of this thread's row, 16-byte stores (conflict-free: lane-contiguous)
#pragma unroll
        for (int kc = 0; kc < 16; ++kc) {
            float v[4], hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * kc + j;
                v[j] = k < Q ? z0[k].x * s : (k < 2 * Q ? z0[k - Q].y * s : 0.f);
                hi[j] = tf32_rn(v[j]);
                lo[j] = tf32_rn(v[j] - hi[j]);
            }
            *reinterpret_cast<float4*>(row_hi + kc * 2048) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(row_lo + kc * 2048) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const uint32_t sa = term == 1 ? sa_lo : sa_hi, sb = term == 2 ? sb_lo : sb_hi;
#pragma unroll
                for (int ks = 0; ks < KD / 8; ++ks)
                    mma_tf32(tbase, smem_desc(sa + ks * 2 * 2048, 2048, 128), smem_desc(sb + ks * 2 * 1024, 1024, 128),
                             (term | ks) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smbar) : "memory");
        }
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smbar), "r"(phase) : "memory");
        } while (!ok);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- results: this thread's row of D (64 FP32) out of tensor memory
#pragma unroll
        for (int c0 = 0; c0 < ND; c0 += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(mine + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) res[c0 + j] = __uint_as_float(r[j]);
        }
        acc_chk += res[1] + res[33];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D is overwritten by the next tile's first MMA
        s += 1.0f + (res[1] + res[32 + 7] + res[28]) * 1e-30f;
    }
    // outputs of the last iteration: re at n = k, im at n = 32 + k
#pragma unroll
    for (int k = 0; k < Q; ++k) y[(size_t)col * Q + k] = gnss::mk(res[k], res[32 + k]);
    if (acc_chk == 1.2345e-30f) y[0].x = acc_chk;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(ND) : "memory");
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int ctas = sms * 16, cols = ctas * 128;
    std::vector<cf> hx((size_t)cols * Q);
    srand(6102);
    for (auto& v : hx) { v.x = (float)rand() / RAND_MAX - 0.5f; v.y = (float)rand() / RAND_MAX - 0.5f; }
    // B = real form of W29, entries rounded to TF32 hi + TF32 lo on the host: bmat[m][n][k]
    std::vector<float> hb(2 * ND * KD, 0.f);
    auto tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    for (int n = 0; n < ND; ++n)
        for (int k = 0; k < KD; ++k) {
            double w = 0.0;
            const bool out_im = n >= 32;
            const int kk = out_im ? n - 32 : n;
            if (kk < Q && k < 2 * Q) {
                const bool in_im = k >= Q;
                const int c = in_im ? k - Q : k;
                const double th = 2.0 * M_PI * ((c * kk) % Q) / Q;
                // (zr + i zi)(cos - i sin): re = zr cos + zi sin ; im = zi cos - zr sin
                w = !out_im ? (in_im ? sin(th) : cos(th)) : (in_im ? cos(th) : -sin(th));
            }
            const float hi = tf32((float)w), lo = tf32((float)(w - (double)hi));
            hb[(0 * ND + n) * KD + k] = hi;
            hb[(1 * ND + n) * KD + k] = lo;
        }
    cf *dx, *dy;
    float* db;
    CK(cudaMalloc(&dx, hx.size() * sizeof(cf)));
    CK(cudaMalloc(&dy, hx.size() * sizeof(cf)));
    CK(cudaMalloc(&db, hb.size() * sizeof(float)));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * sizeof(cf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem_a = 2 * 128 * 29 * sizeof(cf), smem_b = 65536 + 32768 + 1024;
    CK(cudaFuncSetAttribute(dft29_ffma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    CK(cudaFuncSetAttribute(dft29_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    // accuracy: one iteration, scale 1, against float64
    std::vector<cf> hy(hx.size());
    auto check = [&](const char* name) {
        double worst = 0.0, big = 0.0;
        for (int col = 0; col < 4096; ++col)
            for (int k = 0; k < Q; ++k) {
                double re = 0, im = 0;
                for (int c = 0; c < Q; ++c) {
                    const double th = -2.0 * M_PI * ((c * k) % Q) / Q;
                    re += hx[(size_t)col * Q + c].x * cos(th) - hx[(size_t)col * Q + c].y * sin(th);
                    im += hx[(size_t)col * Q + c].x * sin(th) + hx[(size_t)col * Q + c].y * cos(th);
                }
                worst = fmax(worst, hypot(hy[(size_t)col * Q + k].x - re, hy[(size_t)col * Q + k].y - im));
                big = fmax(big, hypot(re, im));
            }
        printf("%-5s: max |err| / max |X| = %.3e\n", name, worst / big);
    };
    dft29_ffma<<<ctas, 128, smem_a>>>(dx, dy, 1, 1.0f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hy.data(), dy, hy.size() * sizeof(cf), cudaMemcpyDeviceToHost));
    check("ffma2");
    dft29_tc5<<<ctas, 128, smem_b>>>(dx, dy, db, 1, 1.0f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hy.data(), dy, hy.size() * sizeof(cf), cudaMemcpyDeviceToHost));
    check("tc5x3");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0);
        dft29_ffma<<<ctas, 128, smem_a>>>(dx, dy, iters, 1.0f);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("ffma2: %.3f ms  %.2f G columns/s\n", ms, (double)cols * iters / ms * 1e-6);
        cudaEventRecord(e0);
        dft29_tc5<<<ctas, 128, smem_b>>>(dx, dy, db, iters, 1.0f);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("tc5x3: %.3f ms  %.2f G columns/s  (24 tcgen05.mma M128 N64 K8 per 128 columns)\n", ms, (double)cols * iters / ms * 1e-6);
    }
    return 0;
}
