// Softmax attention for the short sequences of this path (text: L = 20, causal; ViT-B/32: L = 50), head dim 64:
// one CTA of 4 warps per (sample, head), the whole head in shared memory (bf16, zero padded to 64 rows), warp w owns
// query rows 16w..16w+15.  Products on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in / fp32 accumulate),
// scores / probabilities / softmax in registers with quad shuffles, P and dS exchanged through shared memory for the
// transposed (key-side) products of the backward pass.
//
// Why not tcgen05 here: a 50x50x64 problem is 0.6 MFLOP -- three orders of magnitude below one 128xN UMMA tile pipeline's
// set-up cost; the CTA-wide TMEM/TMA machinery is kept for the dense GEMMs (gemm_sm100.cu) as the north star states.
//
// Replaces nn.MultiheadAttention's core (CLIP/clip/model.py:366-386; causal mask :537-543): softmax(q k^T / 8 + M) v.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.h"

namespace {

constexpr int AP = 72;          // padded smem row (bf16): 144 B, conflict-free for ldmatrix
constexpr int TILE = 64 * AP;   // one 64-row operand

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// L x 64 slice of a row-major bf16 matrix -> smem [64][AP], rows >= L zero.
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ base, long row_stride, int L, __nv_bfloat16* dst) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int l = i >> 3, v = i & 7;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (l < L) r = __ldg(reinterpret_cast<const uint4*>(base + l * row_stride) + v);
        *reinterpret_cast<uint4*>(dst + l * AP + v * 8) = r;
    }
}

// Probabilities of the warp's 16 query rows: p[j][0..1] = row r0+g, keys 8j+2t,+1 ; p[j][2..3] = row r0+g+8.
__device__ __forceinline__ void scores_softmax(const __nv_bfloat16* sQ, const __nv_bfloat16* sK, int r0, int L, int NT, int causal,
                                               float (&p)[8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j][0] = p[j][1] = p[j][2] = p[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        ldsm_x4(a, sQ + (r0 + (lane & 15)) * AP + kk * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < NT) {
                uint32_t b[2];
                ldsm_x2(b, sK + (8 * j + (lane & 7)) * AP + kk * 16 + ((lane >> 3) & 1) * 8);
                mma16816(p[j], a, b);
            }
        }
    }
    const int ra = r0 + g, rb = ra + 8;
    float ma = -CUDART_INF_F, mb = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < NT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 8 * j + 2 * t + (i & 1), r = (i < 2) ? ra : rb;
                const bool dead = c >= L || (causal && c > r);
                p[j][i] = dead ? -CUDART_INF_F : p[j][i] * 0.125f;     // 1/sqrt(64)
            }
            ma = fmaxf(ma, fmaxf(p[j][0], p[j][1]));
            mb = fmaxf(mb, fmaxf(p[j][2], p[j][3]));
        }
    }
    ma = quad_max(ma);
    mb = quad_max(mb);
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < NT) {
            p[j][0] = __expf(p[j][0] - ma); p[j][1] = __expf(p[j][1] - ma);
            p[j][2] = __expf(p[j][2] - mb); p[j][3] = __expf(p[j][3] - mb);
            sa += p[j][0] + p[j][1];
            sb += p[j][2] + p[j][3];
        }
    }
    sa = 1.f / quad_sum(sa);
    sb = 1.f / quad_sum(sb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < NT) { p[j][0] *= sa; p[j][1] *= sa; p[j][2] *= sb; p[j][3] *= sb; }
    }
}

// acc[jd] (16 rows x 64) += X[16 x keys] . M[keys x 64] with X given as C-fragments x[j] (two 8-key tiles per k-step) and M
// row-major [key][64] in smem.
__device__ __forceinline__ void rows_times_matrix(const float (&x)[8][4], const __nv_bfloat16* sM, int NT, int KS, float (&acc)[8][4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (ks < KS) {
            uint32_t a[4];
            a[0] = pack2(x[2 * ks][0], x[2 * ks][1]);
            a[1] = pack2(x[2 * ks][2], x[2 * ks][3]);
            const bool hi = (2 * ks + 1) < NT;
            a[2] = hi ? pack2(x[2 * ks + 1][0], x[2 * ks + 1][1]) : 0u;
            a[3] = hi ? pack2(x[2 * ks + 1][2], x[2 * ks + 1][3]) : 0u;
#pragma unroll
            for (int jd = 0; jd < 8; ++jd) {
                uint32_t b[2];
                ldsm_x2_t(b, sM + (16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * AP + jd * 8);
                mma16816(acc[jd], a, b);
            }
        }
    }
}

__device__ __forceinline__ void store_rows(const float (&acc)[8][4], __nv_bfloat16* dst, long row_stride, int r0, int L) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int jd = 0; jd < 8; ++jd) {
        if (r0 + g < L) *reinterpret_cast<uint32_t*>(dst + (r0 + g) * row_stride + jd * 8 + 2 * t) = pack2(acc[jd][0], acc[jd][1]);
        if (r0 + g + 8 < L) *reinterpret_cast<uint32_t*>(dst + (r0 + g + 8) * row_stride + jd * 8 + 2 * t) = pack2(acc[jd][2], acc[jd][3]);
    }
}

__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int L,
                                                           int heads, int causal) {
    __shared__ __align__(16) __nv_bfloat16 sm[3 * TILE];
    __nv_bfloat16 *sQ = sm, *sK = sm + TILE, *sV = sm + 2 * TILE;
    const int n = blockIdx.x / heads, h = blockIdx.x % heads, D = heads * 64;
    const long row0 = static_cast<long>(n) * L;
    const __nv_bfloat16* base = qkv + row0 * 3 * D + h * 64;
    load_tile(base, 3L * D, L, sQ);
    load_tile(base + D, 3L * D, L, sK);
    load_tile(base + 2 * D, 3L * D, L, sV);
    __syncthreads();
    const int r0 = (threadIdx.x >> 5) * 16;
    if (r0 >= L) return;
    const int NT = (L + 7) >> 3, KS = (L + 15) >> 4;
    float p[8][4], o[8][4];
    scores_softmax(sQ, sK, r0, L, NT, causal, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    rows_times_matrix(p, sV, NT, KS, o);
    store_rows(o, out + row0 * D + h * 64, D, r0, L);
}

__global__ void __launch_bounds__(128) attn_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                           __nv_bfloat16* __restrict__ dqkv, int L, int heads, int causal) {
    extern __shared__ __align__(16) __nv_bfloat16 sm[];
    __nv_bfloat16 *sQ = sm, *sK = sm + TILE, *sV = sm + 2 * TILE, *sG = sm + 3 * TILE, *sP = sm + 4 * TILE, *sD = sm + 5 * TILE;
    const int n = blockIdx.x / heads, h = blockIdx.x % heads, D = heads * 64;
    const long row0 = static_cast<long>(n) * L;
    const __nv_bfloat16* base = qkv + row0 * 3 * D + h * 64;
    load_tile(base, 3L * D, L, sQ);
    load_tile(base + D, 3L * D, L, sK);
    load_tile(base + 2 * D, 3L * D, L, sV);
    load_tile(dout + row0 * D + h * 64, D, L, sG);
    for (int i = threadIdx.x; i < 2 * 64 * (AP / 8); i += blockDim.x) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);   // sP, sD
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = (threadIdx.x >> 5) * 16;
    const int NT = (L + 7) >> 3, KS = (L + 15) >> 4;
    __nv_bfloat16* dbase = dqkv + row0 * 3 * D + h * 64;
    if (r0 < L) {
        float p[8][4], dp[8][4];
        scores_softmax(sQ, sK, r0, L, NT, causal, p);
        // dP = dO V^T
#pragma unroll
        for (int j = 0; j < 8; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t a[4];
            ldsm_x4(a, sG + (r0 + (lane & 15)) * AP + kk * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < NT) {
                    uint32_t b[2];
                    ldsm_x2(b, sV + (8 * j + (lane & 7)) * AP + kk * 16 + ((lane >> 3) & 1) * 8);
                    mma16816(dp[j], a, b);
                }
            }
        }
        float da = 0.f, db = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < NT) {
                da += p[j][0] * dp[j][0] + p[j][1] * dp[j][1];
                db += p[j][2] * dp[j][2] + p[j][3] * dp[j][3];
            }
        }
        da = quad_sum(da);
        db = quad_sum(db);
        // dS (already carrying the 1/8 of the scaled scores); P and dS to smem for the key-side products
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < NT) {
                dp[j][0] = p[j][0] * (dp[j][0] - da) * 0.125f; dp[j][1] = p[j][1] * (dp[j][1] - da) * 0.125f;
                dp[j][2] = p[j][2] * (dp[j][2] - db) * 0.125f; dp[j][3] = p[j][3] * (dp[j][3] - db) * 0.125f;
                const int c = 8 * j + 2 * t;
                *reinterpret_cast<uint32_t*>(sP + (r0 + g) * AP + c) = pack2(p[j][0], p[j][1]);
                *reinterpret_cast<uint32_t*>(sP + (r0 + g + 8) * AP + c) = pack2(p[j][2], p[j][3]);
                *reinterpret_cast<uint32_t*>(sD + (r0 + g) * AP + c) = pack2(dp[j][0], dp[j][1]);
                *reinterpret_cast<uint32_t*>(sD + (r0 + g + 8) * AP + c) = pack2(dp[j][2], dp[j][3]);
            }
        }
        float dq[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
        rows_times_matrix(dp, sK, NT, KS, dq);                     // dQ = dS K
        store_rows(dq, dbase, 3L * D, r0, L);
    }
    __syncthreads();
    if (r0 < L) {                                                  // key rows r0..r0+15: dV = P^T dO, dK = dS^T Q
        float dv[8][4], dk[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                           // 16 queries per step
            if (ks < KS) {
                const int i = lane >> 3;
                const int off = (16 * ks + (lane & 7) + (i >> 1) * 8) * AP + r0 + (i & 1) * 8;
                uint32_t aP[4], aD[4];
                ldsm_x4_t(aP, sP + off);
                ldsm_x4_t(aD, sD + off);
#pragma unroll
                for (int jd = 0; jd < 8; ++jd) {
                    const int boff = (16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * AP + jd * 8;
                    uint32_t b[2];
                    ldsm_x2_t(b, sG + boff);
                    mma16816(dv[jd], aP, b);
                    ldsm_x2_t(b, sQ + boff);
                    mma16816(dk[jd], aD, b);
                }
            }
        }
        store_rows(dk, dbase + D, 3L * D, r0, L);
        store_rows(dv, dbase + 2 * D, 3L * D, r0, L);
    }
}

}  // namespace

namespace tris {

int attn_fwd_mma(const void* qkv, void* out, int n, int L, int heads, int causal, cudaStream_t stream) {
    attn_fwd_mma_kernel<<<n * heads, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), L,
                                                      heads, causal);
    TRIS_LAUNCH_OK("attn_fwd_mma_kernel");
    return TRIS_OK;
}

int attn_bwd_mma(const void* qkv, const void* dout, void* dqkv, int n, int L, int heads, int causal, cudaStream_t stream) {
    static bool attr = false;
    const int smem = 6 * TILE * 2;
    if (!attr) { TRIS_CUDA_OK(cudaFuncSetAttribute(attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
    attn_bwd_mma_kernel<<<n * heads, 128, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
                                                         reinterpret_cast<__nv_bfloat16*>(dqkv), L, heads, causal);
    TRIS_LAUNCH_OK("attn_bwd_mma_kernel");
    return TRIS_OK;
}

}  // namespace tris
