// Does a K-major SWIZZLE_128B UMMA operand work when its start address is shifted by whole 128-byte rows (not 1024-aligned)
// and its 8-row groups sit at a stride (SBO) that is not a multiple of 1024 bytes?  (halo re-use for 3x3 convs: one
// (th+2) x (tw+2) pixel box in smem, nine tap views = nine start addresses.)
// A_logical[R][c], R < 192 rows of 64 bf16, written to smem exactly as TMA SWIZZLE_128B would (16-byte chunk index ^ (R & 7)).
// D = A_view * B^T with B = 64x64 identity  =>  D[m][n] = A_logical[row(m)][n],  row(m) = R0 + (m / 8) * PITCH + (m % 8).
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"

__global__ void k(int R0, int PITCH, int use_base_offset, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // 256 rows x 128 B
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 32768);         // 64 rows x 128 B
    for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
        const int R = i / 64, c = i % 64;
        const float v = (float)((R * 7 + c * 3) % 251) - 125.f;                // exactly representable in bf16
        A[R * 64 + (((c >> 3) ^ (R & 7)) << 3) + (c & 7)] = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, c = i % 64;
        B[n * 64 + (((c >> 3) ^ (n & 7)) << 3) + (c & 7)] = __float2bfloat16(n == c ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 64); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32 && ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(smem) + R0 * 128, sb = ptx::smem_u32(smem) + 32768;
        const uint32_t idesc = ptx::umma_idesc(1u, 0u, 0u, 128, 64);
        uint64_t da0 = ptx::umma_smem_desc_sw128(0, 0u, PITCH * 128);
        if (use_base_offset) da0 |= (uint64_t)((sa >> 7) & 7) << 49;
        const uint64_t db0 = ptx::umma_smem_desc_sw128(0, 0u, 1024);
        for (int ks = 0; ks < 4; ++ks)
            ptx::umma_f16(tmem, da0 | (uint64_t)(((sa >> 4) + ks * 2) & 0x3FFF), db0 | (uint64_t)(((sb >> 4) + ks * 2) & 0x3FFF), idesc, ks > 0);
        ptx::umma_commit(ptx::smem_u32(&bar));
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t raw[32];
    for (int ch = 0; ch < 2; ++ch) {
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, raw);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + ch * 32 + i] = __uint_as_float(raw[i]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

int main() {
    float* d; cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int bo = 0; bo < 2; ++bo)
        for (int pitch : {8, 10, 18})
            for (int R0 : {0, 1, 3, 8, 11}) {
                cudaMemset(d, 0, sizeof(h));
                k<<<1, 128, 64 * 1024>>>(R0, pitch, bo, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 64; ++n) {
                        const int R = R0 + (m / 8) * pitch + (m % 8);
                        const float want = (float)((R * 7 + n * 3) % 251) - 125.f;
                        if (h[m * 64 + n] != want) ++bad;
                    }
                printf("base_offset_field=%d pitch=%2d rows R0=%2d : %s (%d of 8192 wrong) %s\n", bo, pitch, R0, bad ? "MISMATCH" : "exact", bad,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
