// Micro-benchmark: issue rate of tcgen05.mma kind::f16 M=128, N in {32,64,128,256}, K=16 from smem descriptors,
// (a) all into ONE accumulator, (b) round-robin over R accumulators, (c) with a commit every 4 MMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I tris_b200/csrc -o mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"

__global__ void k(int N, int R, int iters, int commit_every, int kmajor, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar[0]), 1); ptx::mbar_init(ptx::smem_u32(&bar[1]), 1); ptx::fence_mbar_init(); }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = ptx::smem_u32(smem), sb = sa + 16384;
        const uint32_t idesc = ptx::umma_idesc(1u, kmajor ? 0u : 1u, kmajor ? 0u : 1u, 128, N);
        const uint64_t da0 = ptx::umma_smem_desc_sw128(0, kmajor ? 0u : 8192u, 1024), db0 = da0;
        const uint32_t kstep = (kmajor ? 32u : 2048u) >> 4;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (i % R) * N;
            const int ks = i & 3;
            ptx::umma_f16(d, da0 | (uint64_t)((sa >> 4) + ks * kstep), db0 | (uint64_t)((sb >> 4) + ks * kstep), idesc, i >= R);
            if (commit_every && (i % commit_every) == commit_every - 1) ptx::umma_commit(ptx::smem_u32(&bar[1]));
        }
        const long long t1 = clock64();
        ptx::umma_commit(ptx::smem_u32(&bar[0]));
        ptx::mbar_wait(ptx::smem_u32(&bar[0]), 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2048;
    for (int km = 1; km >= 0; --km)
    for (int N : {32, 64, 128, 256})
        for (int R : {1, 2, 4})
            for (int ce : {0, 4}) {
                if (R * N > 512) continue;
                k<<<1, 128, 100 * 1024>>>(N, R, iters, ce, km, d);
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                printf("%s N=%3d accs=%d commit_every=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d) %s\n", km ? "K-major " : "MN-major", N, R, ce,
                       (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
