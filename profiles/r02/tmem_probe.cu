// tmem_probe.cu -- tensor memory (TMEM, 256 KB per SM on sm_100a) as per-thread private storage:
// tcgen05.alloc / tcgen05.st / tcgen05.ld round trip (correctness) and the read rate of a 64-column strip per
// thread against the same strip in shared memory.  The search kernel keeps each thread's slice of the
// conjugate code spectrum there (it is re-read for every one of the K blocks of a row).  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int T = 128, COLS = 64;

__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot, int warp) {
    if (warp == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, int warp) {
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(a)),
                 "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float& a, float& b, float& c, float& d) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2); d = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: TMEM strip, MODE 1: shared-memory strip ([COLS][T] floats, conflict-free)
template <int MODE>
__global__ void __launch_bounds__(T) probe(float* out, int iters, int* bad) {
    __shared__ uint32_t slot;
    extern __shared__ float strip[];
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t base = 0, mine = 0;
    if (MODE == 0) {
        base = tmem_alloc(&slot, warp);
        mine = base + ((uint32_t)(32 * warp) << 16);           // lane field: this warp's quarter of the 128 lanes
        for (int c = 0; c < COLS; c += 4) tmem_st4(mine + c, tid + 0.25f * c, tid + 0.25f * (c + 1), tid + 0.25f * (c + 2), tid + 0.25f * (c + 3));
        tmem_wait_st();
    } else {
        for (int c = 0; c < COLS; ++c) strip[c * T + tid] = tid + 0.25f * c;
        __syncthreads();
    }
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < COLS; c += 16) {
            float v[16];
            if (MODE == 0) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) tmem_ld4(mine + c + q, v[q], v[q + 1], v[q + 2], v[q + 3]);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = strip[(c + q) * T + tid];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) acc = fmaf(v[q], 1.0001f, acc);
        }
    }
    // correctness of the round trip (one pass)
    if (MODE == 0) {
        for (int c = 0; c < COLS; c += 4) {
            float a, b, cc, d;
            tmem_ld4(mine + c, a, b, cc, d);
            tmem_wait_ld();
            if (a != tid + 0.25f * c || b != tid + 0.25f * (c + 1) || cc != tid + 0.25f * (c + 2) || d != tid + 0.25f * (c + 3)) atomicAdd(bad, 1);
        }
        tmem_free(base, warp);
    }
    out[blockIdx.x * T + tid] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; int* bad;
    const int grid = sms * 4, iters = 2000;
    CK(cudaMalloc(&d, grid * T * sizeof(float)));
    CK(cudaMalloc(&bad, sizeof(int)));
    CK(cudaMemset(bad, 0, sizeof(int)));
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<grid, T>>>(d, iters, bad); else probe<1><<<grid, T, COLS * T * sizeof(float)>>>(d, iters, bad);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double bytes = (double)grid * T * iters * COLS * 4.0;
        printf("%-28s %8.3f ms  %7.1f GB/s aggregate  %6.1f B/clk/SM (1.965 GHz, 4 CTAs x 128 threads per SM)\n",
               mode == 0 ? "TMEM strip (tcgen05.ld x4)" : "shared-memory strip (LDS.32)", best, bytes / best * 1e-6, bytes / (best * 1e-3) / 1.965e9 / sms);
    }
    int hb = 0;
    CK(cudaMemcpy(&hb, bad, sizeof hb, cudaMemcpyDeviceToHost));
    printf("round-trip mismatches: %d\n", hb);
    return hb != 0;
}
