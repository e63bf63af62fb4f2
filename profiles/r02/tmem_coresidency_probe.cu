// tmem_coresidency_probe.cu -- can several CTAs per SM that each hold a tcgen05.alloc'ed strip of tensor memory be
// co-resident?  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for any kernel that contains tcgen05.alloc,
// which is what pinned the search kernel's TMEM variant to 9 CTA groups under a cooperative launch
// (tmem_cc_variant.txt).  This probe launches PER_SM x #SM CTAs with the search kernel's footprint (128 threads,
// 44 kB of shared memory, 64 TMEM columns each) in an ordinary launch; every CTA allocates its columns, checks in and
// waits (bounded: 20 ms) until all have checked in.  If they all meet, the hardware co-schedules them and only the
// occupancy query is conservative.  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_coresidency_probe tmem_coresidency_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int T = 128;

template <int COLS>
__global__ void __launch_bounds__(T, 4) probe(unsigned* arrived, unsigned total, int* met, int* sm_of, long long budget_clk) {
    __shared__ uint32_t slot;
    extern __shared__ float pad[];
    const int tid = threadIdx.x, warp = tid >> 5;
    pad[tid] = (float)tid;                                     // touch the dynamic shared memory
    if (warp == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    // one round trip through the strip, so that the allocation is really used
    const uint32_t mine = base + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(mine), "r"(tid) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t back;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(back) : "r"(mine) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sm_of[blockIdx.x] = (int)smid;
        atomicAdd(arrived, 1u);
        const long long t0 = clock64();
        int ok = 0;
        while (clock64() - t0 < budget_clk) {
            if (atomicAdd(arrived, 0u) >= total) { ok = 1; break; }
            __nanosleep(500);
        }
        met[blockIdx.x] = ok && back == (uint32_t)tid;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

template <int COLS>
static void run(int per_sm, int sms, size_t smem) {
    const int n = per_sm * sms;
    unsigned* arrived; int *met, *sm_of;
    CK(cudaMalloc(&arrived, 4)); CK(cudaMalloc(&met, n * 4)); CK(cudaMalloc(&sm_of, n * 4));
    CK(cudaMemset(arrived, 0, 4)); CK(cudaMemset(met, 0, n * 4)); CK(cudaMemset(sm_of, 0xff, n * 4));
    CK(cudaFuncSetAttribute(probe<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<COLS>, T, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<COLS><<<n, T, smem>>>(arrived, (unsigned)n, met, sm_of, 40000000ll /* ~20 ms */);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<int> hm(n), hs(n);
    CK(cudaMemcpy(hm.data(), met, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs.data(), sm_of, n * 4, cudaMemcpyDeviceToHost));
    int ok = 0; std::vector<int> cnt(sms + 64, 0); int mx = 0;
    for (int i = 0; i < n; ++i) { ok += hm[i]; if (hs[i] >= 0 && hs[i] < (int)cnt.size()) mx = cnt[hs[i]] + 1 > mx ? ++cnt[hs[i]] : (++cnt[hs[i]], mx); }
    mx = 0; for (int c : cnt) mx = c > mx ? c : mx;
    printf("cols %3d  CTAs/SM asked %d  grid %4d  occupancy API says %d/SM  all met: %s (%d of %d)  max CTAs seen on one SM %d  kernel %.3f ms\n",
           COLS, per_sm, n, occ, ok == n ? "YES" : "no", ok, n, mx, ms);
    cudaFree(arrived); cudaFree(met); cudaFree(sm_of);
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 44 * 1024;
    run<64>(1, sms, smem);
    run<64>(2, sms, smem);
    run<64>(4, sms, smem);
    run<128>(4, sms, smem);
    run<32>(4, sms, smem);
    run<256>(2, sms, smem);
    return 0;
}
