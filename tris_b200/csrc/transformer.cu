// Non-GEMM pieces of the CLIP transformer blocks (text towers, width 512 / 8 heads / L=20 causal; aux ViT-B/32,
// width 768 / 12 heads / L=50): token embedding, LayerNorm fwd/bwd, small-sequence multi-head attention fwd/bwd held
// entirely in shared memory, row gather / scatter-add, column sums (bias gradients).
//
// Replaces nn.Embedding + positional add (CLIP/clip/model.py:553-554), LayerNorm (:352-358), nn.MultiheadAttention's
// softmax(QK^T/sqrt(d)+mask)V (:369,381) and the EOT gather (:562) of the reference; the linears run in gemm_sm100.cu.
#include <cuda_bf16.h>
#include <math_constants.h>

#include <stdlib.h>

#include "common.h"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------- embedding
// one block per sentence: x[n,l,:] = E[ids[n,l]] + P[l]; eot[n] = n*L + argmax_l ids[n,l] (first maximum)
template <bool XF32>   // XF32: the residual stream starts in fp32 (as the reference's autocast keeps it)
__global__ void embed_fwd_kernel(const int* __restrict__ ids, const float* __restrict__ E, const float* __restrict__ P,
                                 void* __restrict__ xv, int* __restrict__ eot, int L, int D) {
    const int n = blockIdx.x;
    if (threadIdx.x == 0 && eot != nullptr) {
        int best = 0, bv = ids[n * L];
        for (int l = 1; l < L; ++l) {
            const int v = ids[n * L + l];
            if (v > bv) { bv = v; best = l; }
        }
        eot[n] = n * L + best;
    }
    for (int i = threadIdx.x; i < L * D / 4; i += blockDim.x) {
        const int l = (i * 4) / D, d = (i * 4) % D;
        const int tok = ids[n * L + l];
        const float4 e = __ldg(reinterpret_cast<const float4*>(E + static_cast<long>(tok) * D + d));
        const float4 p = __ldg(reinterpret_cast<const float4*>(P + l * D + d));
        const long off = (static_cast<long>(n) * L + l) * D + d;
        if (XF32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(xv) + off) = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
        } else {
            __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(xv) + off);
            o[0] = __floats2bfloat162_rn(e.x + p.x, e.y + p.y);
            o[1] = __floats2bfloat162_rn(e.z + p.z, e.w + p.w);
        }
    }
}

// dE[ids] += dx ; dP[l] += sum_n dx, without atomics (bit-reproducible):
//  * positional part: one thread per (l, d) adds the N samples in order;
//  * token part: one CTA per token position i = n*L + l.  The CTA whose position is the FIRST occurrence of its token id owns
//    that embedding row: it adds the rows of every later occurrence in position order (SOT / EOT occur in every sentence).
__global__ void embed_bwd_pos_kernel(const __nv_bfloat16* __restrict__ dx, float* __restrict__ dP, int N, int L, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * D) return;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += __bfloat162float(dx[static_cast<long>(n) * L * D + i]);
    dP[i] += s;
}
__global__ void __launch_bounds__(128) embed_bwd_tok_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ dx,
                                                            float* __restrict__ dE, int T, int D) {
    extern __shared__ int s_ids[];          // [T] token ids, then [T] positions sharing this CTA's id (owner CTAs only)
    __shared__ int s_cnt;
    const int me = blockIdx.x;
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_ids[i] = ids[i];
    __syncthreads();
    const int tok = s_ids[me];
    int earlier = 0;
    for (int j = threadIdx.x; j < me; j += blockDim.x) earlier |= (s_ids[j] == tok);
    if (__syncthreads_or(earlier)) return;  // an earlier position owns this token's row
    if (threadIdx.x == 0) {                 // ordered list of the later occurrences (shared-memory scan, owners only)
        int cnt = 0;
        for (int j = me; j < T; ++j)
            if (s_ids[j] == tok) s_ids[T + cnt++] = j;
        s_cnt = cnt;
    }
    __syncthreads();
    const int cnt = s_cnt;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int k0 = 0; k0 < cnt; k0 += 8) {       // eight rows in flight, added in position order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = k0 + u < cnt ? __bfloat162float(dx[static_cast<long>(s_ids[T + k0 + u]) * D + d]) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        dE[static_cast<long>(tok) * D + d] += s;
    }
}

// ---------------------------------------------------------------------------------------- LayerNorm (one warp per row)
// 128-bit loads: lane handles vectors lane, lane+32, ... of 8 bf16 (D % 256 == 0: 512 -> 2, 768 -> 3, 1024 -> 4 per lane)
struct V8 { float v[8]; };
__device__ __forceinline__ V8 ldv8(const __nv_bfloat16* p) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    V8 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); o.v[2 * j] = f.x; o.v[2 * j + 1] = f.y; }
    return o;
}
__device__ __forceinline__ void stv8(__nv_bfloat16* p, const V8& x) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(x.v[2 * j], x.v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ V8 ldf8(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    V8 o;
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
    return o;
}

// row vectors of the residual stream: bf16, or fp32 (F32) -- element offset `off`
template <bool F32>
__device__ __forceinline__ V8 ldx8(const void* base, long off) {
    if (F32) return ldf8(reinterpret_cast<const float*>(base) + off);
    return ldv8(reinterpret_cast<const __nv_bfloat16*>(base) + off);
}
template <bool F32>
__device__ __forceinline__ void stx8(void* base, long off, const V8& x) {
    if (F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
        p[0] = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
        p[1] = make_float4(x.v[4], x.v[5], x.v[6], x.v[7]);
    } else {
        stv8(reinterpret_cast<__nv_bfloat16*>(base) + off, x);
    }
}

template <int NV, bool XF32, bool YF32>   // vectors per lane = D / 256
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, void* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int rows, int D, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long rbase = static_cast<long>(row) * D;
    V8 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = ldx8<XF32>(x, rbase + (i * 32 + lane) * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i].v[j];
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i].v[j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        const V8 g = ldf8(gamma + c), b = ldf8(beta + c);
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - mean) * rstd * g.v[j] + b.v[j];
        stx8<YF32>(y, rbase + c, o);
    }
    if (lane == 0 && mean_out != nullptr) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// dx = (add ? add : 0) + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma ; dgamma += dy*xhat ; dbeta += dy
template <int NV, bool XF32>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const void* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ add,
                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ ws,
                                                            int rows, int D) {
    extern __shared__ float sm[];   // [2*D] block partials of dgamma/dbeta
    const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool want_param = ws != nullptr;
    if (want_param) {
        for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
    }
    V8 gam[NV], pg[NV], pb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        gam[i] = ldf8(gamma + (i * 32 + lane) * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) pg[i].v[j] = pb[i].v[j] = 0.f;
    }
    for (int row = blockIdx.x * warps + wid; row < rows; row += gridDim.x * warps) {
        const long base = static_cast<long>(row) * D;
        const float mean = mean_in[row], rstd = rstd_in[row];
        V8 g[NV], xh[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 8;
            const V8 d = ldv8(dy + base + c), xv = ldx8<XF32>(x, base + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[i].v[j] = (xv.v[j] - mean) * rstd;
                g[i].v[j] = d.v[j] * gam[i].v[j];
                s1 += g[i].v[j];
                s2 = fmaf(g[i].v[j], xh[i].v[j], s2);
                pg[i].v[j] = fmaf(d.v[j], xh[i].v[j], pg[i].v[j]);
                pb[i].v[j] += d.v[j];
            }
        }
        s1 = warp_sum(s1) / D;
        s2 = warp_sum(s2) / D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 8;
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = rstd * (g[i].v[j] - s1 - xh[i].v[j] * s2);
            if (add != nullptr) {
                const V8 a = ldv8(add + base + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] += a.v[j];
            }
            stv8(dx + base + c, o);
        }
    }
    if (want_param) {
        // fixed order: the warps add their partials to the block sums one after the other; the block sums become row
        // blockIdx.x of the two planes ws[2][gridDim.x][D] (dgamma | dbeta), reduced in row order by the caller's queue
        for (int w = 0; w < warps; ++w) {
            if (wid == w) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sm[(i * 32 + lane) * 8 + j] += pg[i].v[j];
                        sm[D + (i * 32 + lane) * 8 + j] += pb[i].v[j];
                    }
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < D; i += blockDim.x) {
            ws[static_cast<long>(blockIdx.x) * D + i] = sm[i];
            ws[(static_cast<long>(gridDim.x) + blockIdx.x) * D + i] = sm[D + i];
        }
    }
}

// ---------------------------------------------------------------------------------------- attention (L <= 64, head dim 64)
constexpr int HD = 64;
constexpr int HDP = HD + 1;

// one block (128 threads) per (sequence, head).  qkv: [N*L, 3*D] bf16 (q | k | v), out: [N*L, D].
// All contractions are 4x4 register tiles (0.5 shared-memory loads per FMA); softmax rows by warp shuffles.
__device__ __forceinline__ void softmax_rows(float* s, float* ds, int L, bool bwd) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int a = wid; a < L; a += 4) {
        float m = -CUDART_INF_F;
        for (int b = lane; b < L; b += 32) m = fmaxf(m, s[a * (L + 1) + b]);
        m = warp_max(m);
        float sum = 0.f;
        for (int b = lane; b < L; b += 32) {
            const float e = __expf(s[a * (L + 1) + b] - m);
            s[a * (L + 1) + b] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        if (!bwd) {
            for (int b = lane; b < L; b += 32) s[a * (L + 1) + b] *= inv;
        } else {
            float delta = 0.f;
            for (int b = lane; b < L; b += 32) {
                const float pr = s[a * (L + 1) + b] * inv;
                s[a * (L + 1) + b] = pr;
                delta += pr * ds[a * (L + 1) + b];
            }
            delta = warp_sum(delta);
            for (int b = lane; b < L; b += 32) ds[a * (L + 1) + b] = s[a * (L + 1) + b] * (ds[a * (L + 1) + b] - delta);
        }
    }
}

// One head's L x 64 slice of a row-major bf16 matrix -> fp32 smem [L][HDP], 128-bit loads (8 vectors per row).
__device__ __forceinline__ void load_head_rows(const __nv_bfloat16* __restrict__ base, long row_stride, int L, float* dst, float scale) {
    for (int i = threadIdx.x; i < L * 8; i += blockDim.x) {
        const int l = i >> 3, vv = i & 7;
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(base + l * row_stride) + vv);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
        float* d = dst + l * HDP + vv * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            d[2 * j] = f.x * scale;
            d[2 * j + 1] = f.y * scale;
        }
    }
}

template <bool BWD>
__device__ __forceinline__ void score_tiles(const float* q, const float* k, const float* go, const float* v, float* s, float* ds,
                                            int L, int causal) {
    const int nt = (L + 3) >> 2;
    for (int id = threadIdx.x; id < nt * nt; id += blockDim.x) {
        const int a0 = (id / nt) * 4, b0 = (id % nt) * 4;
        int ar[4], br[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ar[i] = min(a0 + i, L - 1) * HDP; br[i] = min(b0 + i, L - 1) * HDP; }
        float acc[4][4], dp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = dp[i][jj] = 0.f;
        for (int d = 0; d < HD; ++d) {
            float qa[4], kb[4], ga[4], vb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { qa[i] = q[ar[i] + d]; kb[i] = k[br[i] + d]; }
            if (BWD) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { ga[i] = go[ar[i] + d]; vb[i] = v[br[i] + d]; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    acc[i][jj] = fmaf(qa[i], kb[jj], acc[i][jj]);
                    if (BWD) dp[i][jj] = fmaf(ga[i], vb[jj], dp[i][jj]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int a = a0 + i, b = b0 + jj;
                if (a < L && b < L) {
                    s[a * (L + 1) + b] = (causal && b > a) ? -CUDART_INF_F : acc[i][jj];
                    if (BWD) ds[a * (L + 1) + b] = (causal && b > a) ? 0.f : dp[i][jj];
                }
            }
    }
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                       int L, int heads, int causal) {
    extern __shared__ float sm[];
    float* q = sm;
    float* k = q + L * HDP;
    float* v = k + L * HDP;
    float* s = v + L * HDP;   // [L][L+1]
    const int n = blockIdx.x / heads, h = blockIdx.x % heads;
    const int D = heads * HD;
    const long row0 = static_cast<long>(n) * L;
    const __nv_bfloat16* base = qkv + row0 * 3 * D + h * HD;
    load_head_rows(base, 3L * D, L, q, 0.125f);   // 1/sqrt(64)
    load_head_rows(base + D, 3L * D, L, k, 1.f);
    load_head_rows(base + 2 * D, 3L * D, L, v, 1.f);
    __syncthreads();
    score_tiles<false>(q, k, nullptr, nullptr, s, nullptr, L, causal);
    __syncthreads();
    softmax_rows(s, nullptr, L, false);
    __syncthreads();
    const int nta = (L + 3) >> 2;
    for (int id = threadIdx.x; id < nta * 16; id += blockDim.x) {
        const int a0 = (id >> 4) * 4, d0 = (id & 15) * 4;
        int ar[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ar[i] = min(a0 + i, L - 1) * (L + 1);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
        for (int b = 0; b < L; ++b) {
            float pa[4], vb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { pa[i] = s[ar[i] + b]; vb[i] = v[b * HDP + d0 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(pa[i], vb[jj], acc[i][jj]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (a0 + i < L) {
                __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(out + (row0 + a0 + i) * D + h * HD + d0);
                o[0] = __floats2bfloat162_rn(acc[i][0], acc[i][1]);
                o[1] = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
            }
    }
}

// dqkv from dout, recomputing the probabilities.
__global__ void __launch_bounds__(128) attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                       __nv_bfloat16* __restrict__ dqkv, int L, int heads, int causal) {
    extern __shared__ float sm[];
    float* q = sm;
    float* k = q + L * HDP;
    float* v = k + L * HDP;
    float* go = v + L * HDP;
    float* s = go + L * HDP;          // P  [L][L+1]
    float* ds = s + L * (L + 1);      // dS [L][L+1]
    const int n = blockIdx.x / heads, h = blockIdx.x % heads;
    const int D = heads * HD;
    const long row0 = static_cast<long>(n) * L;
    const __nv_bfloat16* base = qkv + row0 * 3 * D + h * HD;
    load_head_rows(base, 3L * D, L, q, 0.125f);
    load_head_rows(base + D, 3L * D, L, k, 1.f);
    load_head_rows(base + 2 * D, 3L * D, L, v, 1.f);
    load_head_rows(dout + row0 * D + h * HD, D, L, go, 1.f);
    __syncthreads();
    score_tiles<true>(q, k, go, v, s, ds, L, causal);
    __syncthreads();
    softmax_rows(s, ds, L, true);
    __syncthreads();
    const int nta = (L + 3) >> 2;
    for (int id = threadIdx.x; id < nta * 16; id += blockDim.x) {
        const int a0 = (id >> 4) * 4, d0 = (id & 15) * 4;
        int ac[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ac[i] = min(a0 + i, L - 1);
        float dq[4][4], dk[4][4], dv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dq[i][jj] = dk[i][jj] = dv[i][jj] = 0.f;
        for (int b = 0; b < L; ++b) {
            float dsa[4], dst[4], pt[4], kb[4], qb[4], gb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dsa[i] = ds[ac[i] * (L + 1) + b];
                dst[i] = ds[b * (L + 1) + ac[i]];
                pt[i] = s[b * (L + 1) + ac[i]];
                kb[i] = k[b * HDP + d0 + i];
                qb[i] = q[b * HDP + d0 + i];      // q already carries the 1/8 scale
                gb[i] = go[b * HDP + d0 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    dq[i][jj] = fmaf(dsa[i], kb[jj], dq[i][jj]);
                    dk[i][jj] = fmaf(dst[i], qb[jj], dk[i][jj]);
                    dv[i][jj] = fmaf(pt[i], gb[jj], dv[i][jj]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (a0 + i < L) {
                __nv_bfloat16* r = dqkv + (row0 + a0 + i) * 3 * D + h * HD + d0;
                __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(r);
                o[0] = __floats2bfloat162_rn(dq[i][0] * 0.125f, dq[i][1] * 0.125f);
                o[1] = __floats2bfloat162_rn(dq[i][2] * 0.125f, dq[i][3] * 0.125f);
                o = reinterpret_cast<__nv_bfloat162*>(r + D);
                o[0] = __floats2bfloat162_rn(dk[i][0], dk[i][1]);
                o[1] = __floats2bfloat162_rn(dk[i][2], dk[i][3]);
                o = reinterpret_cast<__nv_bfloat162*>(r + 2 * D);
                o[0] = __floats2bfloat162_rn(dv[i][0], dv[i][1]);
                o[1] = __floats2bfloat162_rn(dv[i][2], dv[i][3]);
            }
    }
}

// ---------------------------------------------------------------------------------------- gather / scatter / colsum
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                   int rows, int D) {
    const int vecs = D >> 3;
    for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < static_cast<long>(rows) * vecs;
         v += static_cast<long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(v / vecs), c = static_cast<int>(v % vecs);
        reinterpret_cast<uint4*>(out)[v] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long>(idx[r]) * D) + c);
    }
}
// out (pre-zeroed by this kernel's first phase is not possible) -> caller zeroes; out[idx[r]] = src[r] (unique idx)
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                    int rows, int D) {
    const int vecs = D >> 3;
    for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < static_cast<long>(rows) * vecs;
         v += static_cast<long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(v / vecs), c = static_cast<int>(v % vecs);
        reinterpret_cast<uint4*>(out + static_cast<long>(idx[r]) * D)[c] = __ldg(reinterpret_cast<const uint4*>(src) + v);
    }
}

// ws[chunk][c] = sum over the chunk's rows of x[r, c]   (x bf16 [rows, N]); grid = (ceil(N/64), row_chunks), block = (64, 4).
// Plain stores: the caller's queue adds the chunk rows in order into the bias gradient (bit-reproducible).
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int rows, int N) {
    __shared__ float part[4][64];
    const int c = blockIdx.x * 64 + threadIdx.x;
    float s = 0.f;
    if (c < N)
        for (int r = blockIdx.y * 4 + threadIdx.y; r < rows; r += gridDim.y * 4) s += __bfloat162float(x[static_cast<long>(r) * N + c]);
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N)
        out[static_cast<long>(blockIdx.y) * N + c] = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
}

// ViT token assembly: tok[n,0,:] = cls + pos[0]; tok[n,1+p,:] = patch[n*P+p,:] + pos[1+p]
__global__ void vit_assemble_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                    __nv_bfloat16* __restrict__ tok, int N, int T, int D) {
    const long total = static_cast<long>(N) * T * D;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i % D);
        const long r = i / D;
        const int t = static_cast<int>(r % T);
        const long n = r / T;
        const float base = (t == 0) ? cls[d] : __bfloat162float(patch[(n * (T - 1) + (t - 1)) * D + d]);
        tok[i] = __float2bfloat16(base + pos[t * D + d]);
    }
}

}  // namespace

namespace tris {
int attn_fwd_mma(const void* qkv, void* out, int n, int L, int heads, int causal, cudaStream_t stream);
int attn_bwd_mma(const void* qkv, const void* dout, void* dqkv, int n, int L, int heads, int causal, cudaStream_t stream);
}
// TRIS_ATTN_FP32=1 selects the register-tiled fp32 CUDA-core kernels below instead of the tensor-core ones (attn_small.cu).
static bool use_fp32_attention() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TRIS_ATTN_FP32"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int NV>
static void launch_ln_fwd(dim3 grid, int threads, cudaStream_t st, int x_f32, int y_f32, const void* x, const float* gamma,
                          const float* beta, void* y, float* mean, float* rstd, int rows, int D, float eps) {
    if (x_f32 && y_f32) layernorm_fwd_kernel<NV, true, true><<<grid, threads, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
    else if (x_f32) layernorm_fwd_kernel<NV, true, false><<<grid, threads, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
    else if (y_f32) layernorm_fwd_kernel<NV, false, true><<<grid, threads, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
    else layernorm_fwd_kernel<NV, false, false><<<grid, threads, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
}

template <int NV>
static void launch_ln_bwd(int grid, int threads, size_t smb, cudaStream_t st, int x_f32, const __nv_bfloat16* dy, const void* x,
                          const float* gamma, const float* mean, const float* rstd, const __nv_bfloat16* add, __nv_bfloat16* dx,
                          float* ws, int rows, int D) {
    if (x_f32) layernorm_bwd_kernel<NV, true><<<grid, threads, smb, st>>>(dy, x, gamma, mean, rstd, add, dx, ws, rows, D);
    else layernorm_bwd_kernel<NV, false><<<grid, threads, smb, st>>>(dy, x, gamma, mean, rstd, add, dx, ws, rows, D);
}

extern "C" {

int tris_embed_fwd(const int* ids, const float* E, const float* P, void* x, int* eot, int n, int L, int D, int x_f32,
                   tris_stream_t stream) {
    if (D % 4) return tris::fail(TRIS_ERR_SHAPE, "tris_embed_fwd: D %% 4");
    if (x_f32) embed_fwd_kernel<true><<<n, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, E, P, x, eot, L, D);
    else embed_fwd_kernel<false><<<n, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, E, P, x, eot, L, D);
    TRIS_LAUNCH_OK("embed_fwd_kernel");
    return TRIS_OK;
}

int tris_embed_bwd(const int* ids, const void* dx, float* dE, float* dP, int n, int L, int D, tris_stream_t stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* dxp = reinterpret_cast<const __nv_bfloat16*>(dx);
    const int T = n * L;
    if (2 * T * sizeof(int) > 48 * 1024) return tris::fail(TRIS_ERR_SHAPE, "tris_embed_bwd: %d tokens exceed the shared-memory id table", T);
    embed_bwd_pos_kernel<<<(L * D + 255) / 256, 256, 0, st>>>(dxp, dP, n, L, D);
    TRIS_LAUNCH_OK("embed_bwd_pos_kernel");
    embed_bwd_tok_kernel<<<T, 128, 2 * T * sizeof(int), st>>>(ids, dxp, dE, T, D);
    TRIS_LAUNCH_OK("embed_bwd_tok_kernel");
    return TRIS_OK;
}

int tris_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int rows, int D,
                       float eps, int x_f32, int y_f32, tris_stream_t stream) {
    if (D % 256 || D > 1024) return tris::fail(TRIS_ERR_SHAPE, "tris_layernorm_fwd: D=%d must be a multiple of 256, <= 1024", D);
    const int warps = 8;
    const dim3 grid((rows + warps - 1) / warps);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (D / 256) {
        case 1: launch_ln_fwd<1>(grid, warps * 32, st, x_f32, y_f32, x, gamma, beta, y, mean, rstd, rows, D, eps); break;
        case 2: launch_ln_fwd<2>(grid, warps * 32, st, x_f32, y_f32, x, gamma, beta, y, mean, rstd, rows, D, eps); break;
        case 3: launch_ln_fwd<3>(grid, warps * 32, st, x_f32, y_f32, x, gamma, beta, y, mean, rstd, rows, D, eps); break;
        default: launch_ln_fwd<4>(grid, warps * 32, st, x_f32, y_f32, x, gamma, beta, y, mean, rstd, rows, D, eps); break;
    }
    TRIS_LAUNCH_OK("layernorm_fwd_kernel");
    return TRIS_OK;
}

int tris_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* add,
                       void* dx, float* ws, int ws_rows, int rows, int D, int x_f32, tris_stream_t stream) {
    if (D % 256 || D > 1024) return tris::fail(TRIS_ERR_SHAPE, "tris_layernorm_bwd: D=%d must be a multiple of 256, <= 1024", D);
    const int warps = 8;
    int grid = (rows + warps - 1) / warps;
    if (ws != nullptr) {
        // parameter gradients: exactly ws_rows CTAs, each writes row blockIdx.x of ws[2][ws_rows][D]
        if (ws_rows < 1 || ws_rows > grid) return tris::fail(TRIS_ERR_SHAPE, "tris_layernorm_bwd: ws_rows %d not in 1..%d", ws_rows, grid);
        grid = ws_rows;
    } else if (grid > 4 * tris::sm_count()) {
        grid = 4 * tris::sm_count();
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16 *dyp = reinterpret_cast<const __nv_bfloat16*>(dy), *ap = reinterpret_cast<const __nv_bfloat16*>(add);
    __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
    const size_t smb = 2 * D * sizeof(float);
    switch (D / 256) {
        case 1: launch_ln_bwd<1>(grid, warps * 32, smb, st, x_f32, dyp, x, gamma, mean, rstd, ap, dxp, ws, rows, D); break;
        case 2: launch_ln_bwd<2>(grid, warps * 32, smb, st, x_f32, dyp, x, gamma, mean, rstd, ap, dxp, ws, rows, D); break;
        case 3: launch_ln_bwd<3>(grid, warps * 32, smb, st, x_f32, dyp, x, gamma, mean, rstd, ap, dxp, ws, rows, D); break;
        default: launch_ln_bwd<4>(grid, warps * 32, smb, st, x_f32, dyp, x, gamma, mean, rstd, ap, dxp, ws, rows, D); break;
    }
    TRIS_LAUNCH_OK("layernorm_bwd_kernel");
    return TRIS_OK;
}

static size_t attn_smem(int L, bool bwd) {
    return static_cast<size_t>((bwd ? 4 : 3) * L * HDP + (bwd ? 2 : 1) * L * (L + 1)) * sizeof(float);
}

int tris_attn_fwd(const void* qkv, void* out, int n, int L, int heads, int causal, tris_stream_t stream) {
    if (L >= 1 && L <= 64 && !use_fp32_attention()) return tris::attn_fwd_mma(qkv, out, n, L, heads, causal, reinterpret_cast<cudaStream_t>(stream));
    if (L > 64 || L < 1) return tris::fail(TRIS_ERR_SHAPE, "tris_attn_fwd: L=%d must be in 1..64", L);
    static bool attr = false;
    if (!attr) { TRIS_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
    attn_fwd_kernel<<<n * heads, 128, attn_smem(L, false), reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), L, heads, causal);
    TRIS_LAUNCH_OK("attn_fwd_kernel");
    return TRIS_OK;
}

int tris_attn_bwd(const void* qkv, const void* dout, void* dqkv, int n, int L, int heads, int causal, tris_stream_t stream) {
    if (L >= 1 && L <= 64 && !use_fp32_attention()) return tris::attn_bwd_mma(qkv, dout, dqkv, n, L, heads, causal, reinterpret_cast<cudaStream_t>(stream));
    if (L > 64 || L < 1) return tris::fail(TRIS_ERR_SHAPE, "tris_attn_bwd: L=%d must be in 1..64", L);
    static bool attr = false;
    if (!attr) { TRIS_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024)); attr = true; }
    attn_bwd_kernel<<<n * heads, 128, attn_smem(L, true), reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
        reinterpret_cast<__nv_bfloat16*>(dqkv), L, heads, causal);
    TRIS_LAUNCH_OK("attn_bwd_kernel");
    return TRIS_OK;
}

int tris_gather_rows(const void* x, const int* idx, void* out, int rows, int D, tris_stream_t stream) {
    if (D % 8) return tris::fail(TRIS_ERR_SHAPE, "tris_gather_rows: D %% 8");
    const long total = static_cast<long>(rows) * (D / 8);
    gather_rows_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), idx, reinterpret_cast<__nv_bfloat16*>(out), rows, D);
    TRIS_LAUNCH_OK("gather_rows_kernel");
    return TRIS_OK;
}

int tris_scatter_rows(const void* src, const int* idx, void* out, int rows, int D, tris_stream_t stream) {
    if (D % 8) return tris::fail(TRIS_ERR_SHAPE, "tris_scatter_rows: D %% 8");
    const long total = static_cast<long>(rows) * (D / 8);
    scatter_rows_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(src), idx, reinterpret_cast<__nv_bfloat16*>(out), rows, D);
    TRIS_LAUNCH_OK("scatter_rows_kernel");
    return TRIS_OK;
}

int tris_colsum(const void* x, float* ws, int chunks, int rows, int N, tris_stream_t stream) {
    if (chunks < 1 || chunks > 65535) return tris::fail(TRIS_ERR_SHAPE, "tris_colsum: chunks %d", chunks);
    dim3 grid((N + 63) / 64, chunks);
    colsum_kernel<<<grid, dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ws, rows, N);
    TRIS_LAUNCH_OK("colsum_kernel");
    return TRIS_OK;
}

int tris_vit_assemble(const void* patch, const float* cls, const float* pos, void* tok, int n, int T, int D, tris_stream_t stream) {
    const long total = static_cast<long>(n) * T * D;
    int grid = static_cast<int>((total + 255) / 256);
    if (grid > 8 * tris::sm_count()) grid = 8 * tris::sm_count();
    vit_assemble_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(patch), cls, pos, reinterpret_cast<__nv_bfloat16*>(tok), n, T, D);
    TRIS_LAUNCH_OK("vit_assemble_kernel");
    return TRIS_OK;
}

}  // extern "C"
