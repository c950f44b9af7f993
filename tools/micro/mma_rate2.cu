// Micro-benchmark v2: cost model of tcgen05.mma kind::f16 on one SM.
//   mode 0: SS (A, B from smem), M=128      mode 1: SS, M=64      mode 2: TS (A from TMEM), M=128
//   issuers: 1 or 2 threads (different warps) issuing concurrently into different accumulators
//   commit_every: tcgen05.commit to a ring of 8 mbarriers every k MMAs (0 = never)
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void k(int mode, int N, int issuers, int iters, int commit_every, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[20];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 20; ++i) ptx::mbar_init(ptx::smem_u32(&bar[i]), 1); ptx::fence_mbar_init(); }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int me = warp;   // issuer id = warp id
    if ((threadIdx.x & 31) == 0 && me < issuers) {
        const uint32_t sa = ptx::smem_u32(smem) + me * 49152, sb = sa + 16384;
        const int M = mode == 1 ? 64 : 128;
        const uint32_t idesc = ptx::umma_idesc(1u, 0u, 0u, M, N);
        const uint64_t d0 = ptx::umma_smem_desc_sw128(0, 0u, 1024);
        const uint32_t dcol = tmem + me * 256, acol = tmem + 448;     // A operand (TS) in the last 64 columns: 128 x 16 bf16 = 8 columns
        const unsigned long long g0 = gtime();
        const long long t0 = clock64();
        int ring = 0;
        for (int i = 0; i < iters; ++i) {
            const int ks = i & 3;
            if (mode == 2) umma_ts(dcol, acol + ks * 8, d0 | (uint64_t)((sb >> 4) + ks * 2), idesc, i > 0);
            else ptx::umma_f16(dcol, d0 | (uint64_t)((sa >> 4) + ks * 2), d0 | (uint64_t)((sb >> 4) + ks * 2), idesc, i > 0);
            if (commit_every && (i % commit_every) == commit_every - 1) { ptx::umma_commit(ptx::smem_u32(&bar[2 + me * 8 + ring])); ring = (ring + 1) & 7; }
        }
        const long long t1 = clock64();
        ptx::umma_commit(ptx::smem_u32(&bar[me]));
        ptx::mbar_wait(ptx::smem_u32(&bar[me]), 0);
        const long long t2 = clock64();
        const unsigned long long g1 = gtime();
        out[me * 3 + 0] = t1 - t0; out[me * 3 + 1] = t2 - t0; out[me * 3 + 2] = (long long)(g1 - g0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    const int iters = 4096;
    const char* names[3] = {"SS M=128", "SS M=64 ", "TS M=128"};
    for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {32, 64, 128, 256})
            for (int issuers : {1, 2})
                for (int ce : {0, 4, 16}) {
                    if (rep == 0 && !(mode == 0 && N == 256 && issuers == 1 && ce == 0)) continue;   // warm-up pass
                    cudaMemset(d, 0, 64);
                    k<<<1, 128, 120 * 1024>>>(mode, N, issuers, iters, ce, d);
                    long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
                    cudaError_t e = cudaGetLastError();
                    if (rep == 0) continue;
                    printf("%s N=%3d issuers=%d commit_every=%2d : issue %.1f  complete %.1f cyc/mma, %.1f ns/mma (%.2f GHz)%s\n", names[mode], N, issuers, ce,
                           (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[1] / (double)h[2], e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
