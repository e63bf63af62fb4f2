// tc_dft29_probe.cu -- is a DFT-as-GEMM radix-29 stage worth moving to the tensor cores?
//
// BASELINE.json's north_star: "Tensor cores are used only if a DFT-as-GEMM radix stage is shown ... to beat the
// CUDA-core butterflies."  Pass 1 of the search kernel is a 29-point DFT per column (840 FFMA in the
// symmetric-pair form of gnss_radix.h).  As a GEMM it is  C = Wc * a  (15 x 15 cosines) and  S = Ws * b
// (14 x 14 sines) on a_j = x_j + x_{29-j}, b_j = x_j - x_{29-j}, real and imaginary parts as separate
// right-hand sides.  FP32 accuracy (peak/SNR within 1e-4, top-2 ties at 2e-5) rules out plain TF32; the
// error-compensated 3xTF32 split (hi*hi + lo*hi + hi*lo) keeps ~2^-21.  This probe times, on register-resident
// data,
//   A: dft_odd<29> (the product's butterflies), one column per thread;
//   B: mma.sync.m16n8k8 tf32, 3-way split, 8 columns per warp step (24 MMAs), fragments formed in registers;
// both store their 29 outputs per column to shared memory as the product's pass 1 does, and checks B against a
// float64 DFT.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../assignment-for-aae6102_gnss-sdr_b200/csrc
//                      tc_dft29_probe.cu -o tc_dft29_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "gnss_radix.h"

using gnss::cf;
constexpr int Q = 29, H = 14;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- A: CUDA-core butterflies
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
        for (int c = 0; c < Q; ++c) out[c * 128 + threadIdx.x] = v[c];      // conflict-free: lane-contiguous
        s += 1.0f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < Q; ++c) y[(size_t)col * Q + c] = out[c * 128 + threadIdx.x];   // last iteration's result
}

// ---------------------------------------------------------------- B: tensor cores, 3xTF32
__device__ __forceinline__ unsigned tf32_hi(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += W * B with W = Whi + Wlo, B = Bhi + Blo (drops lo*lo)
__device__ __forceinline__ void mma3(float (&d)[4], const unsigned (&whi)[4], const unsigned (&wlo)[4], float b0, float b1) {
    const unsigned h0 = tf32_hi(b0), h1 = tf32_hi(b1);
    const unsigned l0 = tf32_hi(b0 - __uint_as_float(h0)), l1 = tf32_hi(b1 - __uint_as_float(h1));
    mma_tf32(d, wlo, h0, h1);
    mma_tf32(d, whi, l0, l1);
    mma_tf32(d, whi, h0, h1);
}

// wtab: [mat 0 cos / 1 sin][kstep 2][hi/lo 2][reg 4][lane 32]
__global__ void __launch_bounds__(128, 3) dft29_mma(const cf* __restrict__ x, cf* __restrict__ y, const unsigned* __restrict__ wtab,
                                                    int iters, float s0) {
    __shared__ cf out[4 * 8 * 30 * 4];                    // [warp][col-in-group 8][k 29 (+1)] x 4 groups per iteration
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    unsigned W[2][2][2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                for (int r = 0; r < 4; ++r) W[m][s][hl][r] = wtab[(((m * 2 + s) * 2 + hl) * 4 + r) * 32 + lane];
    // this thread's inputs: column gid of each of the warp's 4 groups of 8 columns, j = tig + 4*u (u = 0..3)
    const int col_base = (blockIdx.x * 4 + warp) * 32;    // 32 columns per warp per iteration = 4 groups
    cf xin[4][4][2];                                      // [group][u][x_j, x_{29-j}]
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = tig + 4 * u;
            const size_t base = (size_t)(col_base + g * 8 + gid) * Q;
            xin[g][u][0] = (j < 15) ? x[base + j] : gnss::mk(0.f, 0.f);
            xin[g][u][1] = (j >= 1 && j < 15) ? x[base + Q - j] : gnss::mk(0.f, 0.f);
        }
    float s = s0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float a_re[4], a_im[4], b_re[4], b_im[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const cf p = gnss::mk(xin[g][u][0].x * s, xin[g][u][0].y * s), q = gnss::mk(xin[g][u][1].x * s, xin[g][u][1].y * s);
                a_re[u] = p.x + q.x; a_im[u] = p.y + q.y;               // j = 0: q = 0 -> a_0 = x_0; j = 15: zeros
                b_re[u] = p.x - q.x; b_im[u] = p.y - q.y;
            }
            if (tig == 0) { b_re[0] = 0.f; b_im[0] = 0.f; }             // b_0 does not exist
            float Cre[4] = {0, 0, 0, 0}, Cim[4] = {0, 0, 0, 0}, Sre[4] = {0, 0, 0, 0}, Sim[4] = {0, 0, 0, 0};
#pragma unroll
            for (int st = 0; st < 2; ++st) {                            // k-step: j = tig + 8 st, tig + 4 + 8 st
                mma3(Cre, W[0][st][0], W[0][st][1], a_re[2 * st], a_re[2 * st + 1]);
                mma3(Cim, W[0][st][0], W[0][st][1], a_im[2 * st], a_im[2 * st + 1]);
                mma3(Sre, W[1][st][0], W[1][st][1], b_re[2 * st], b_re[2 * st + 1]);
                mma3(Sim, W[1][st][0], W[1][st][1], b_im[2 * st], b_im[2 * st + 1]);
            }
            // accumulator (row = gid + 8 h, col = 2 tig + e) -> X_k = C - iS, X_{29-k} = C + iS, k = gid + 8 h
            cf* o = out + (size_t)((warp * 4 + g) * 8) * 30;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = gid + 8 * h, c = 2 * tig + e, r = 2 * h + e;
                    if (k <= H) {
                        o[c * 30 + k] = gnss::mk(Cre[r] + Sim[r], Cim[r] - Sre[r]);
                        if (k >= 1) o[c * 30 + Q - k] = gnss::mk(Cre[r] - Sim[r], Cim[r] + Sre[r]);
                    }
                }
        }
        s += 1.0f;
    }
    __syncwarp();
    for (int g = 0; g < 4; ++g)
        for (int e = lane; e < 8 * Q; e += 32) {
            const int c = e / Q, k = e % Q;
            y[(size_t)(col_base + g * 8 + c) * Q + k] = out[(size_t)((warp * 4 + g) * 8 + c) * 30 + k];
        }
}

static unsigned host_tf32(float v) {                      // round-to-nearest-away at 10 mantissa bits, like cvt.rna
    unsigned u;
    memcpy(&u, &v, 4);
    u += 0x1000u;
    return u & 0xffffe000u;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    const int blocks = 148 * 4 * 4, cols = blocks * 128;
    std::vector<cf> hx((size_t)cols * Q), hy((size_t)cols * Q);
    srand(6102);
    for (auto& v : hx) { v.x = (rand() / (float)RAND_MAX - 0.5f) * 200.f; v.y = (rand() / (float)RAND_MAX - 0.5f) * 200.f; }
    // twiddle fragments
    std::vector<unsigned> wt(2 * 2 * 2 * 4 * 32);
    for (int m = 0; m < 2; ++m)
        for (int s = 0; s < 2; ++s)
            for (int r = 0; r < 4; ++r)
                for (int lane = 0; lane < 32; ++lane) {
                    const int gid = lane >> 2, tig = lane & 3;
                    const int row = gid + 8 * (r & 1), colj = tig + 4 * (r >> 1) + 8 * s;    // a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4)
                    double w = 0.0;
                    if (m == 0) { if (row <= H && colj <= H) w = colj == 0 ? 1.0 : cos(2.0 * M_PI * ((row * colj) % Q) / Q); }
                    else        { if (row >= 1 && row <= H && colj >= 1 && colj <= H) w = sin(2.0 * M_PI * ((row * colj) % Q) / Q); }
                    const unsigned hi = host_tf32((float)w);
                    float hif; memcpy(&hif, &hi, 4);
                    const unsigned lo = host_tf32((float)(w - (double)hif));
                    wt[(((m * 2 + s) * 2 + 0) * 4 + r) * 32 + lane] = hi;
                    wt[(((m * 2 + s) * 2 + 1) * 4 + r) * 32 + lane] = lo;
                }
    cf *dx, *dy; unsigned* dw;
    CK(cudaMalloc(&dx, hx.size() * sizeof(cf))); CK(cudaMalloc(&dy, hy.size() * sizeof(cf))); CK(cudaMalloc(&dw, wt.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * sizeof(cf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem_a = 2 * 128 * 29 * sizeof(cf);
    CK(cudaFuncSetAttribute(dft29_ffma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto check = [&](const char* name) {
        cudaMemcpy(hy.data(), dy, hy.size() * sizeof(cf), cudaMemcpyDeviceToHost);
        double worst = 0.0, scale = 0.0;
        for (int col = 0; col < 64; ++col)
            for (int k = 0; k < Q; ++k) {
                double re = 0, im = 0;
                for (int n = 0; n < Q; ++n) {
                    const double ang = -2.0 * M_PI * ((n * k) % Q) / Q;
                    re += hx[(size_t)col * Q + n].x * cos(ang) - hx[(size_t)col * Q + n].y * sin(ang);
                    im += hx[(size_t)col * Q + n].x * sin(ang) + hx[(size_t)col * Q + n].y * cos(ang);
                }
                worst = fmax(worst, fmax(fabs(hy[(size_t)col * Q + k].x - re), fabs(hy[(size_t)col * Q + k].y - im)));
                scale = fmax(scale, fmax(fabs(re), fabs(im)));
            }
        printf("%s: max |err| / max |X| = %.3e\n", name, worst / scale);
    };
    float ms;
    // accuracy (one iteration, scale 1)
    dft29_ffma<<<blocks, 128, smem_a>>>(dx, dy, 1, 1.0f); CK(cudaDeviceSynchronize()); check("ffma ");
    CK(cudaMemset(dy, 0, hy.size() * sizeof(cf)));
    dft29_mma<<<blocks, 128>>>(dx, dy, dw, 1, 1.0f); CK(cudaDeviceSynchronize()); check("mma3 ");
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); dft29_ffma<<<blocks, 128, smem_a>>>(dx, dy, iters, 1.0f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("ffma : %.3f ms  %.2f G columns/s  (%.1f TFLOP/s at 1680 flop/column)\n", ms, (double)cols * iters / ms / 1e6, (double)cols * iters * 1680 / ms / 1e9);
        CK(cudaEventRecord(e0)); dft29_mma<<<blocks, 128>>>(dx, dy, dw, iters, 1.0f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("mma3 : %.3f ms  %.2f G columns/s\n", ms, (double)cols * iters / ms / 1e6);
    }
    return 0;
}
