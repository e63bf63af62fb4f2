// fp32x2_probe.cu -- what do the packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2) buy?
// Measures lane throughput (FMA lanes per clock per SM) of scalar FFMA vs FFMA2 in its operand forms
// (register pair, 32-bit immediate broadcast, scalar-register broadcast, LO_HI swap), FADD2, and the
// issue-slot relief when shared-memory loads are interleaved.  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NACC = 12, ITERS = 2048, T = 256;

template <int MODE>
__global__ void __launch_bounds__(T) probe(float2* out, float2 m, float2 c, int iters) {
    __shared__ float2 sh[T * 2];
    sh[threadIdx.x] = m; sh[threadIdx.x + T] = c;
    __syncthreads();
    float2 a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float s = 0.f; float2 s2 = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) {            // scalar FFMA x2 (reg,reg,reg)
                a[i].x = fmaf(a[i].x, m.x, c.x);
                a[i].y = fmaf(a[i].y, m.y, c.y);
            } else if (MODE == 1) {     // FFMA2 reg pairs
                a[i] = __ffma2_rn(a[i], m, c);
            } else if (MODE == 2) {     // FFMA2 immediate broadcast
                a[i] = __ffma2_rn(a[i], make_float2(0.99993f, 0.99993f), c);
            } else if (MODE == 3) {     // FFMA2 scalar register broadcast
                a[i] = __ffma2_rn(a[i], make_float2(m.x, m.x), c);
            } else if (MODE == 4) {     // FFMA2 with swapped operand (complex-style)
                a[i] = __ffma2_rn(make_float2(a[i].y, a[i].x), m, c);
            } else if (MODE == 5) {     // FADD2
                a[i] = __fadd2_rn(a[i], m);
            } else if (MODE == 6) {     // scalar FADD x2
                a[i].x += m.x; a[i].y += m.y;
            } else if (MODE == 7) {     // scalar FFMA immediate x2
                a[i].x = fmaf(a[i].x, 0.99993f, c.x);
                a[i].y = fmaf(a[i].y, 0.99993f, c.y);
            } else if (MODE == 8) {     // scalar FFMA x2 + one LDS.64 per 4 FMA pairs
                a[i].x = fmaf(a[i].x, m.x, c.x);
                a[i].y = fmaf(a[i].y, m.y, c.y);
                if ((i & 3) == 0) { float2 v = sh[(threadIdx.x + it + i) & (2 * T - 1)]; s += v.x + v.y; }
            } else if (MODE == 9) {     // FFMA2 + one LDS.64 per 4 FFMA2
                a[i] = __ffma2_rn(a[i], m, c);
                if ((i & 3) == 0) { float2 v = sh[(threadIdx.x + it + i) & (2 * T - 1)]; s2 = __fadd2_rn(s2, v); }
            }
        }
    }
    float2 r = make_float2(s + s2.x, s2.y);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * T + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float2* d, int sms, double clk_ghz) {
    const int grid = sms * 8;
    probe<MODE><<<grid, T>>>(d, make_float2(0.99991f, 1.00003f), make_float2(1e-4f, -1e-4f), ITERS);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, T>>>(d, make_float2(0.99991f, 1.00003f), make_float2(1e-4f, -1e-4f), ITERS);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lane_ops = (double)grid * T * ITERS * NACC * 2.0;   // one lane-op = one FMA (or add) on one float
    const double per_clk_sm = lane_ops / (best * 1e-3) / (clk_ghz * 1e9) / sms;
    printf("%-44s %8.3f ms  %7.2f T lane-ops/s  %6.1f lanes/clk/SM (at %.3f GHz)\n", name, best, lane_ops / best * 1e-9,
           per_clk_sm, clk_ghz);
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double ghz = khz * 1e-6;
    float2* d; CK(cudaMalloc(&d, (size_t)sms * 8 * T * sizeof(float2)));
    printf("SMs %d, max clock %.3f GHz\n", sms, ghz);
    run<0>("scalar FFMA (2 per complex lane)", d, sms, ghz);
    run<7>("scalar FFMA immediate", d, sms, ghz);
    run<1>("FFMA2 register pairs", d, sms, ghz);
    run<2>("FFMA2 immediate broadcast", d, sms, ghz);
    run<3>("FFMA2 scalar-register broadcast", d, sms, ghz);
    run<4>("FFMA2 swapped (LO_HI) operand", d, sms, ghz);
    run<6>("scalar FADD", d, sms, ghz);
    run<5>("FADD2", d, sms, ghz);
    run<8>("scalar FFMA + LDS.64 every 4 pairs", d, sms, ghz);
    run<9>("FFMA2 + LDS.64 every 4", d, sms, ghz);
    return 0;
}
