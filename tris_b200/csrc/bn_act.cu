// BatchNorm2d (train/eval) + ReLU + residual add + 2x2 average pool on NHWC bf16 activations: forward apply,
// backward reduction and backward apply.  HBM-bound streaming kernels: 128-bit loads/stores (8 channels per
// thread), per-channel scale/shift staged in shared memory, fp32 math.  All reductions are two-stage with a FIXED order
// (per-CTA partial rows, then an in-order sum in a small finalize kernel): no floating-point atomics, results are
// bit-reproducible run to run.
//
// Replaces nn.BatchNorm2d / ReLU / AvgPool2d / "out += identity" of CLIP/clip/model.py:18-55, :212-232, :255-260.
// Batch statistics (sum, sum of squares) arrive from the conv GEMM epilogue (gemm_sm100.cu).
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.h"

namespace {

constexpr int kThreads = 256;

// Programmatic dependent launch: the tiny finalize kernels and the apply kernels behind them are launched while their
// predecessor still runs (launch latency hidden); griddepcontrol.wait blocks until the predecessor grid has completed and
// its writes are visible.  Measured on the full step: 18.79 ms with, 18.82 ms without -- no gain, so it is OFF unless
// TRIS_PDL=1 (the griddepcontrol instructions are no-ops for normally launched grids).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool use_pdl() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TRIS_PDL"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}
template <class P>
cudaError_t launch_dep(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const P& params) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, params);
}

struct Vec8 {
    float v[8];
};

__device__ __forceinline__ Vec8 ld8(const __nv_bfloat16* p) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        o.v[2 * j] = f.x;
        o.v[2 * j + 1] = f.y;
    }
    return o;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const Vec8& x) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(x.v[2 * j], x.v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}

struct BnBranch {
    const __nv_bfloat16* y;   // conv output [M_in, C]
    const float* stats;       // [2C] sum, sumsq (train) or nullptr
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float* save_mean;         // [C] out (train) / in (backward)
    float* save_invstd;
};

struct ApplyParams {
    BnBranch b0, b1;          // b1.y == nullptr -> single branch
    const __nv_bfloat16* residual;  // [M_out, C] or nullptr
    __nv_bfloat16* out;       // [M_out, C]
    unsigned char* relu_bits; // optional [M_out, C/8]: bit i of byte (row, c/8) = (out[row, c + i] > 0); lets the backward
                              // kernels mask the upstream gradient without re-reading the bf16 output (1/16 of its bytes)
    int n, h, w, c, pool, relu, train;
    int unpair;                  // pooled form only: channel halves are the two images of a pair -> write out as [2n, ho, wo, c/2]
    float count, momentum, eps;
};

struct FinalizeParams;
__device__ __forceinline__ void bn_prologue(const BnBranch& b, const ApplyParams& p, float* s_scale, float* s_shift) {
    for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
        // train: batch mean / invstd were finalized by bn_fwd_finalize_kernel (fixed-order sum of the GEMM's partial rows)
        const float mean = p.train ? __ldcg(b.save_mean + c) : b.running_mean[c];
        const float invstd = p.train ? __ldcg(b.save_invstd + c) : rsqrtf(b.running_var[c] + p.eps);
        const float sc = b.gamma[c] * invstd;
        s_scale[c] = sc;
        s_shift[c] = b.beta[c] - mean * sc;
    }
}


// In-order sum of the per-CTA partial rows parts[nparts][2C] = (sum x | sum x^2) -> batch mean / invstd (+ running stats,
// momentum update with the unbiased variance: nn.BatchNorm2d, CLIP/clip/model.py:18-28).  fold_half > 0: channels c and
// c + fold_half are ONE BatchNorm channel (image-pair-packed stem): their sums are averaged, so that sum / count (count =
// pixels of B/2 images) is the full-batch mean.  blockIdx.y selects the branch.
struct FinalizeParams {
    const float* parts[2];
    float* rm[2]; float* rv[2]; float* mean[2]; float* invstd[2];
    const float* gamma[2]; const float* beta[2];
    float* scale[2]; float* shift[2];     // optional: gamma * invstd and beta - mean * gamma * invstd (recomputed-mask epilogues)
    int nparts, c, fold_half;
    float count, momentum, eps;
};
// Fixed-order sum over the rows of a partial buffer, cooperative: a CTA of 32 x G threads owns 32 consecutive columns
// (threadIdx.x & 31) and G row groups (threadIdx.x >> 5); group g adds rows g, g + G, ... in order, the G group sums are
// then added in order through shared memory.  The order never depends on timing -> bit-reproducible.
constexpr int kFinGroups = 32;     // 148 partial rows / 32 groups: <= 5 loads per thread, one round trip
constexpr int kFinGroupsBwd = 32;  // up to 592 partial rows of the stand-alone reduction kernel
template <int G>
__device__ __forceinline__ float fin_partial(const float* base, int nrows, long row_stride) {
    // up to ten independent loads in flight per thread, then added in row order (rows past the end contribute an exact +0)
    float s = 0.f;
    constexpr int U = 10;      // 592 rows / 32 groups = 18.5 loads per thread: two batches
    for (int r0 = threadIdx.x >> 5; r0 < nrows; r0 += U * G) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * G;
            v[u] = r < nrows ? __ldg(base + r * row_stride) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) s += v[u];
    }
    return s;
}
template <int G>
__device__ __forceinline__ float2 fin_combine2(float2 (*sm)[33], float2 v) {
    __syncthreads();
    sm[threadIdx.x >> 5][threadIdx.x & 31] = v;
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < G; ++g) { const float2 x = sm[g][threadIdx.x & 31]; t.x += x.x; t.y += x.y; }
    return t;
}
template <int G>
__device__ __forceinline__ float fin_combine(float (*sm)[33], float v) {
    __syncthreads();
    sm[threadIdx.x >> 5][threadIdx.x & 31] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) t += sm[g][threadIdx.x & 31];
    return t;
}

__device__ void bn_fwd_finalize_group(const FinalizeParams& p, int br, int group, float2 (*sm)[33]) {
    const int c = group * 32 + (threadIdx.x & 31);
    const int cc = c < p.c ? c : p.c - 1;                       // tail lanes shadow a valid column (no divergence at the barriers)
    const long rs = 2L * p.c;
    const float* parts = p.parts[br];
    // parameters first: their latency hides behind the partial-row loads
    float gam = 0.f, bet = 0.f, rm = 0.f, rv = 0.f;
    const bool writer = c < p.c && threadIdx.x < 32;
    if (writer) {
        if (p.scale[br] != nullptr) { gam = __ldg(p.gamma[br] + c); bet = __ldg(p.beta[br] + c); }
        if (p.rm[br] != nullptr) { rm = p.rm[br][c]; rv = p.rv[br][c]; }
    }
    float s, q;
    if (p.fold_half > 0) {
        const int lo = cc < p.fold_half ? cc : cc - p.fold_half, hi = lo + p.fold_half;
        const float2 p0 = make_float2(fin_partial<kFinGroups>(parts + lo, p.nparts, rs), fin_partial<kFinGroups>(parts + p.c + lo, p.nparts, rs));
        const float2 p1 = make_float2(fin_partial<kFinGroups>(parts + hi, p.nparts, rs), fin_partial<kFinGroups>(parts + p.c + hi, p.nparts, rs));
        const float2 t0 = fin_combine2<kFinGroups>(sm, p0), t1 = fin_combine2<kFinGroups>(sm, p1);
        s = (t0.x + t1.x) * 0.5f;
        q = (t0.y + t1.y) * 0.5f;
    } else {
        const float2 t = fin_combine2<kFinGroups>(sm, make_float2(fin_partial<kFinGroups>(parts + cc, p.nparts, rs),
                                                                  fin_partial<kFinGroups>(parts + p.c + cc, p.nparts, rs)));
        s = t.x;
        q = t.y;
    }
    if (!writer) return;
    const float mean = s / p.count;
    const float var = fmaxf(q / p.count - mean * mean, 0.f);
    const float invstd = rsqrtf(var + p.eps);
    p.mean[br][c] = mean;
    p.invstd[br][c] = invstd;
    if (p.scale[br] != nullptr) {
        const float sc = gam * invstd;
        p.scale[br][c] = sc;
        p.shift[br][c] = bet - mean * sc;
    }
    if (p.rm[br] != nullptr) {
        const float unbiased = var * (p.count / fmaxf(p.count - 1.f, 1.f));
        p.rm[br][c] = (1.f - p.momentum) * rm + p.momentum * mean;
        p.rv[br][c] = (1.f - p.momentum) * rv + p.momentum * unbiased;
    }
}

// (Folding this step into the head of the apply kernel behind an in-kernel counter was measured SLOWER, 20.0 vs 19.0 ms per
// step: every CTA of the apply grid then pays same-address atomics for the ticket / exit protocol.  It stays a tiny kernel.)
__global__ void __launch_bounds__(32 * kFinGroups) bn_fwd_finalize_kernel(const FinalizeParams p) {
    __shared__ float2 sm[kFinGroups][33];
    pdl_wait();                    // the conv GEMM that wrote the partial rows
    pdl_launch_dependents();       // the apply kernel may be set up now (it waits for this grid itself)
    bn_fwd_finalize_group(p, blockIdx.y, blockIdx.x, sm);
}

template <bool UNPAIR>
__device__ __forceinline__ void bn_apply_fwd_body(const ApplyParams& p) {
    extern __shared__ float sm[];
    float* sc0 = sm;
    float* sh0 = sm + p.c;
    float* sc1 = sm + 2 * p.c;
    float* sh1 = sm + 3 * p.c;
    const bool dual = p.b1.y != nullptr;
    if (p.train) pdl_wait();       // the finalize kernel (and, through it, the conv GEMM)
    bn_prologue(p.b0, p, sc0, sh0);
    if (dual) bn_prologue(p.b1, p, sc1, sh1);
    __syncthreads();
    const int vecs = p.c >> 3;
    const int ho = p.h / p.pool, wo = p.w / p.pool;
    const long rows_out = static_cast<long>(p.n) * ho * wo;
    // the grid stride (gridDim * 256 vectors) is a multiple of `vecs` (vecs | 256): a thread keeps its channel group,
    // so scale/shift live in registers and the loop walks rows without divisions.
    const long start = static_cast<long>(blockIdx.x) * kThreads + threadIdx.x;
    const int c0 = static_cast<int>(start % vecs) * 8;
    const long row_step = static_cast<long>(gridDim.x) * kThreads / vecs;
    float a0[8], b0[8], a1[8], b1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a0[i] = sc0[c0 + i]; b0[i] = sh0[c0 + i];
        a1[i] = dual ? sc1[c0 + i] : 0.f; b1[i] = dual ? sh1[c0 + i] : 0.f;
    }
    for (long orow = start / vecs; orow < rows_out; orow += row_step) {
        Vec8 acc;
        long urow = 0;        // UNPAIR: output row in the un-paired layout
        if (p.pool == 1) {
            acc = ld8(p.b0.y + orow * p.c + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(acc.v[i], a0[i], b0[i]);
            if (dual) {
                Vec8 y1 = ld8(p.b1.y + orow * p.c + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += fmaf(y1.v[i], a1[i], b1[i]);
            }
            if (p.residual != nullptr) {
                Vec8 r = ld8(p.residual + orow * p.c + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += r.v[i];
            }
            if (p.relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] = fmaxf(acc.v[i], 0.f);
            }
        } else {
            const int xo = static_cast<int>(orow % wo);
            const long t = orow / wo;
            const int yo = static_cast<int>(t % ho);
            const long ni = t / ho;
            if (UNPAIR) urow = ((2 * ni + (c0 >= (p.c >> 1))) * ho + yo) * wo + xo;   // (pair ni, half j) -> image 2 ni + j
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long irow = (ni * p.h + (2 * yo + dy)) * p.w + (2 * xo + dx);
                    Vec8 y0 = ld8(p.b0.y + irow * p.c + c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float z = fmaf(y0.v[i], a0[i], b0[i]);
                        if (p.relu) z = fmaxf(z, 0.f);
                        acc.v[i] += 0.25f * z;
                    }
                }
        }
        if (UNPAIR) st8(p.out + urow * (p.c >> 1) + (c0 >= (p.c >> 1) ? c0 - (p.c >> 1) : c0), acc);
        else st8(p.out + orow * p.c + c0, acc);
        if (p.relu_bits != nullptr) {
            unsigned b = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) b |= (acc.v[i] > 0.f ? 1u : 0u) << i;
            p.relu_bits[orow * (p.c >> 3) + (c0 >> 3)] = static_cast<unsigned char>(b);
        }
    }
}

__global__ void __launch_bounds__(kThreads) bn_apply_fwd_kernel(const ApplyParams p) { bn_apply_fwd_body<false>(p); }
// pair-packed stem: un-paired output addressing (a few more registers: capped so that 4 CTAs/SM stay resident)
__global__ void __launch_bounds__(kThreads, 4) bn_apply_fwd_unpair_kernel(const ApplyParams p) { bn_apply_fwd_body<true>(p); }

// ------------------------------------------------------------------------------------------------ backward
struct BwdParams {
    const __nv_bfloat16* dout;   // [M_out, C]
    const __nv_bfloat16* out;    // [M_out, C] forward output (relu mask source) or nullptr -> recompute from y0
    const unsigned char* relu_bits;   // [M_out, C/8] sign bits of the forward output (preferred over `out`: 1/16 of the bytes)
    BnBranch b0, b1;             // y, gamma, beta, save_mean, save_invstd are inputs here
    float* dgamma0; float* dbeta0; float* dgamma1; float* dbeta1;   // [C] fp32 gradient buffers (+= by the finalize kernel)
    float* parts;                // [nparts][K][C] per-CTA partial sums: k = 0 sum g, 1 sum g (y0 - mu0), 2 sum g (y1 - mu1)
    float* red;                  // [K][C] finalized: dbeta, dgamma0, dgamma1 -- what the apply kernel reads
    int nparts, kred, fold_half;
    __nv_bfloat16* dy0; __nv_bfloat16* dy1;   // [M_in, C]
    __nv_bfloat16* g_out;        // optional [M_out, C]: masked upstream gradient (identity branch)
    int n, h, w, c, pool, relu;
    int unpair;                  // pooled form only: dout is [2n, ho, wo, c/2] (un-paired images), read as the pair-packed [n, ho, wo, c]
    float count;
};

// Per-channel constants live in shared memory (loaded per use as two LDS.128) instead of ~90 registers, which kept the
// old kernels at 2 CTAs/SM (128 registers/thread, 23 % of the warps) and latency-bound at ~45 % of the HBM rate.
__device__ __forceinline__ Vec8 lds8(const float* s) {
    Vec8 o;
    const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
    return o;
}

// g at output row orow for channel group c0, masked (ReLU) and pool-scaled.  sc/sh: smem scale/shift of branch 0 (mask
// recompute when the forward output was not kept).
template <bool UNPAIR>
__device__ __forceinline__ Vec8 masked_grad(const BwdParams& p, long orow, int c0, const Vec8& y0, const float* s_sc, const float* s_sh) {
    Vec8 g;
    if (UNPAIR) {
        const int half = p.c >> 1, j = c0 >= half;
        const long hw = static_cast<long>(p.h / 2) * (p.w / 2);
        const long ni = orow / hw, rem = orow - ni * hw;
        g = ld8(p.dout + ((2 * ni + j) * hw + rem) * half + (c0 - j * half));
    } else {
        g = ld8(p.dout + orow * p.c + c0);
    }
    if (p.pool == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g.v[i] *= 0.25f;
    }
    if (p.relu) {
        if (p.relu_bits != nullptr) {
            const unsigned b = __ldg(p.relu_bits + orow * (p.c >> 3) + (c0 >> 3));
#pragma unroll
            for (int i = 0; i < 8; ++i) g.v[i] = ((b >> i) & 1u) ? g.v[i] : 0.f;
        } else if (p.out != nullptr) {
            Vec8 o = ld8(p.out + orow * p.c + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) g.v[i] = o.v[i] > 0.f ? g.v[i] : 0.f;
        } else {
            const Vec8 sc = lds8(s_sc + c0), sh = lds8(s_sh + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) g.v[i] = fmaf(y0.v[i], sc.v[i], sh.v[i]) > 0.f ? g.v[i] : 0.f;
        }
    }
    return g;
}

__device__ __forceinline__ long out_row_of(const BwdParams& p, long irow) {
    if (p.pool == 1) return irow;
    const int x = static_cast<int>(irow % p.w);
    const long t = irow / p.w;
    const int y = static_cast<int>(t % p.h);
    const long ni = t / p.h;
    return (ni * (p.h / 2) + (y >> 1)) * (p.w / 2) + (x >> 1);
}

// dbeta = sum g ; dgamma = invstd * sum g (y - mean).   smem: sc0 [C], sh0 [C], mu0 [C], mu1 [C], then the reduction scratch.
template <bool DUAL, bool UNPAIR>
__global__ void __launch_bounds__(kThreads, 4) bn_bwd_reduce_kernel(const BwdParams p) {
    extern __shared__ float sm[];
    float* s_sc = sm;
    float* s_sh = sm + p.c;
    float* s_mu0 = sm + 2 * p.c;
    float* s_mu1 = sm + 3 * p.c;
    float* red = sm + 4 * p.c;   // [kThreads * 8]
    pdl_launch_dependents();     // the finalize kernel may be set up while this grid runs
    for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
        const float invstd = p.b0.save_invstd[c], mean = p.b0.save_mean[c];
        const float sc = p.b0.gamma[c] * invstd;
        s_sc[c] = sc; s_sh[c] = p.b0.beta[c] - mean * sc; s_mu0[c] = mean;
        s_mu1[c] = DUAL ? p.b1.save_mean[c] : 0.f;
    }
    __syncthreads();
    const int vecs = p.c >> 3;
    const long rows = static_cast<long>(p.n) * p.h * p.w;
    // grid stride is a multiple of `vecs` (vecs | kThreads), so each thread keeps one channel group.
    const long start = static_cast<long>(blockIdx.x) * kThreads + threadIdx.x;
    const int c0 = static_cast<int>(start % vecs) * 8;
    const long row_step = static_cast<long>(gridDim.x) * kThreads / vecs;
    float dg0[8], db[8], dg1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dg0[i] = db[i] = dg1[i] = 0.f;
    for (long irow = start / vecs; irow < rows; irow += row_step) {
        const long orow = out_row_of(p, irow);
        const Vec8 y0 = ld8(p.b0.y + irow * p.c + c0);
        Vec8 y1;
        if (DUAL) y1 = ld8(p.b1.y + irow * p.c + c0);
        const Vec8 g = masked_grad<UNPAIR>(p, orow, c0, y0, s_sc, s_sh);
        const Vec8 mu0 = lds8(s_mu0 + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            db[i] += g.v[i];
            dg0[i] = fmaf(g.v[i], y0.v[i] - mu0.v[i], dg0[i]);
        }
        if (DUAL) {
            const Vec8 mu1 = lds8(s_mu1 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dg1[i] = fmaf(g.v[i], y1.v[i] - mu1.v[i], dg1[i]);
        }
    }
    // block reduction over the kThreads/vecs threads that share a channel group; the CTA's totals become row blockIdx.x of
    // parts[nparts][K][C] (plain stores -- the finalize kernel adds the rows in order)
    const int per = kThreads / vecs;   // threads per channel group (>= 1)
    float* row = p.parts + static_cast<long>(blockIdx.x) * p.kred * p.c;
    auto block_sum = [&](float (&x)[8], float* dst) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = x[i];
        __syncthreads();
        for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
            const int g = c >> 3, i = c & 7;
            float s = 0.f;
            for (int k = 0; k < per; ++k) s += red[(g + k * vecs) * 8 + i];
            dst[c] = s;
        }
    };
    block_sum(db, row);
    block_sum(dg0, row + p.c);
    if (DUAL) block_sum(dg1, row + 2 * p.c);
}

// In-order sum of parts[nparts][K][C] -> red[K][C] = (dbeta, dgamma0, dgamma1) and the parameter gradients (+=), for one group
// of 32 channels (one CTA of bn_bwd_finalize_kernel).
__device__ void bn_bwd_finalize_group(const BwdParams& p, int group, float (*sm)[33]) {
    const int c = group * 32 + (threadIdx.x & 31);
    const int cc = c < p.c ? c : p.c - 1;
    const long rs = static_cast<long>(p.kred) * p.c;
    float part[3][2];
    for (int k = 0; k < p.kred; ++k) {       // all loads first (independent), the barriers of the combines afterwards
        const float* base = p.parts + static_cast<long>(k) * p.c;
        if (p.fold_half > 0) {
            const int lo = cc < p.fold_half ? cc : cc - p.fold_half;
            part[k][0] = fin_partial<kFinGroupsBwd>(base + lo, p.nparts, rs);
            part[k][1] = fin_partial<kFinGroupsBwd>(base + lo + p.fold_half, p.nparts, rs);
        } else {
            part[k][0] = fin_partial<kFinGroupsBwd>(base + cc, p.nparts, rs);
            part[k][1] = 0.f;
        }
    }
    for (int k = 0; k < p.kred; ++k) {
        float s = fin_combine<kFinGroupsBwd>(sm, part[k][0]);
        if (p.fold_half > 0)     // pair-packed stem: channels c and c + fold_half are one BatchNorm channel (averaged)
            s = (s + fin_combine<kFinGroupsBwd>(sm, part[k][1])) * 0.5f;
        if (c >= p.c || threadIdx.x >= 32) continue;
        if (k == 1) s *= p.b0.save_invstd[c];
        if (k == 2) s *= p.b1.save_invstd[c];
        p.red[k * p.c + c] = s;
        if (k == 0) { p.dbeta0[c] += s; if (p.kred == 3) p.dbeta1[c] += s; }
        if (k == 1) p.dgamma0[c] += s;
        if (k == 2) p.dgamma1[c] += s;
    }
}

__global__ void __launch_bounds__(32 * kFinGroupsBwd) bn_bwd_finalize_kernel(const BwdParams p) {
    __shared__ float sm[kFinGroupsBwd][33];
    pdl_wait();
    pdl_launch_dependents();
    bn_bwd_finalize_group(p, blockIdx.x, sm);
}

// dy = a g + b y + c per channel with  a = gamma*invstd,  b = -a*invstd*dgamma/N,  c = a*(invstd*dgamma/N*mean - dbeta/N).
// smem: sc0, sh0, b0, c0, (dual) a1, b1, c1  -- each [C].
template <bool DUAL, bool UNPAIR>
__global__ void __launch_bounds__(kThreads, 4) bn_bwd_apply_kernel(const BwdParams p) {
    extern __shared__ float sm[];
    float* s_sc = sm;
    float* s_sh = sm + p.c;
    float* s_b0 = sm + 2 * p.c;
    float* s_c0 = sm + 3 * p.c;
    float* s_a1 = sm + 4 * p.c;
    float* s_b1 = sm + 5 * p.c;
    float* s_c1 = sm + 6 * p.c;
    const float inv_count = 1.f / p.count;
    pdl_wait();                    // the finalize kernel
    for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
        const float invstd = p.b0.save_invstd[c], mean = p.b0.save_mean[c];
        const float sc = p.b0.gamma[c] * invstd;
        const float k = __ldcg(p.red + p.c + c) * inv_count * invstd, m = __ldcg(p.red + c) * inv_count;
        s_sc[c] = sc; s_sh[c] = p.b0.beta[c] - mean * sc;
        s_b0[c] = -sc * k; s_c0[c] = sc * (k * mean - m);
        if (DUAL) {
            const float is1 = p.b1.save_invstd[c], mu1 = p.b1.save_mean[c];
            const float a1 = p.b1.gamma[c] * is1;
            const float k1 = __ldcg(p.red + 2 * p.c + c) * inv_count * is1, m1 = __ldcg(p.red + c) * inv_count;
            s_a1[c] = a1; s_b1[c] = -a1 * k1; s_c1[c] = a1 * (k1 * mu1 - m1);
        }
    }
    __syncthreads();
    const int vecs = p.c >> 3;
    const long rows = static_cast<long>(p.n) * p.h * p.w;
    const long start = static_cast<long>(blockIdx.x) * kThreads + threadIdx.x;
    const int c0 = static_cast<int>(start % vecs) * 8;
    const long row_step = static_cast<long>(gridDim.x) * kThreads / vecs;
    for (long irow = start / vecs; irow < rows; irow += row_step) {
        const long orow = out_row_of(p, irow);
        const Vec8 y0 = ld8(p.b0.y + irow * p.c + c0);
        Vec8 y1;
        if (DUAL) y1 = ld8(p.b1.y + irow * p.c + c0);
        const Vec8 g = masked_grad<UNPAIR>(p, orow, c0, y0, s_sc, s_sh);
        Vec8 d;
        {
            const Vec8 a = lds8(s_sc + c0), b = lds8(s_b0 + c0), c = lds8(s_c0 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) d.v[i] = fmaf(a.v[i], g.v[i], fmaf(b.v[i], y0.v[i], c.v[i]));
        }
        st8(p.dy0 + irow * p.c + c0, d);
        if (DUAL) {
            const Vec8 a = lds8(s_a1 + c0), b = lds8(s_b1 + c0), c = lds8(s_c1 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) d.v[i] = fmaf(a.v[i], g.v[i], fmaf(b.v[i], y1.v[i], c.v[i]));
            st8(p.dy1 + irow * p.c + c0, d);
        }
        if (p.g_out != nullptr) st8(p.g_out + irow * p.c + c0, g);
    }
}

// ------------------------------------------------------------------------------------------------ avg pool 2x2
__global__ void __launch_bounds__(kThreads) avgpool2_fwd_kernel(const __nv_bfloat16* x, __nv_bfloat16* out, int n, int h,
                                                                int w, int c) {
    const int vecs = c >> 3, ho = h / 2, wo = w / 2;
    const long total = static_cast<long>(n) * ho * wo * vecs;
    for (long v = static_cast<long>(blockIdx.x) * kThreads + threadIdx.x; v < total;
         v += static_cast<long>(gridDim.x) * kThreads) {
        const int c0 = static_cast<int>(v % vecs) * 8;
        const long orow = v / vecs;
        const int xo = static_cast<int>(orow % wo);
        const long t = orow / wo;
        const int yo = static_cast<int>(t % ho);
        const long ni = t / ho;
        Vec8 acc;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                Vec8 a = ld8(x + ((ni * h + 2 * yo + dy) * w + 2 * xo + dx) * c + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += 0.25f * a.v[i];
            }
        st8(out + orow * c + c0, acc);
    }
}

// dx[n,h,w,c] = 0.25 * dout[n,h/2,w/2,c] (+ add[n,h,w,c])
__global__ void __launch_bounds__(kThreads) avgpool2_bwd_kernel(const __nv_bfloat16* dout, const __nv_bfloat16* add,
                                                                __nv_bfloat16* dx, int n, int h, int w, int c) {
    const int vecs = c >> 3;
    const long total = static_cast<long>(n) * h * w * vecs;
    for (long v = static_cast<long>(blockIdx.x) * kThreads + threadIdx.x; v < total;
         v += static_cast<long>(gridDim.x) * kThreads) {
        const int c0 = static_cast<int>(v % vecs) * 8;
        const long irow = v / vecs;
        const int x = static_cast<int>(irow % w);
        const long t = irow / w;
        const int y = static_cast<int>(t % h);
        const long ni = t / h;
        Vec8 g = ld8(dout + ((ni * (h / 2) + (y >> 1)) * (w / 2) + (x >> 1)) * c + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) g.v[i] *= 0.25f;
        if (add != nullptr) {
            Vec8 a = ld8(add + irow * c + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) g.v[i] += a.v[i];
        }
        st8(dx + irow * c + c0, g);
    }
}

int grid_for(long total_threads) {
    long blocks = (total_threads + kThreads - 1) / kThreads;
    static int per_sm = 0;   // TRIS_BN_CTAS: CTAs per SM of the streaming kernels; 4 = one resident wave at 64 registers (19.70 vs 19.81 ms at 8)
    if (per_sm == 0) { const char* e = getenv("TRIS_BN_CTAS"); per_sm = (e && atoi(e) >= 1 && atoi(e) <= 16) ? atoi(e) : 4; }
    const long cap = static_cast<long>(tris::sm_count()) * per_sm;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int check_c(int c, const char* who) {
    if (c % 8 || c > 2048 || (kThreads % (c / 8)) != 0)
        return tris::fail(TRIS_ERR_SHAPE, "%s: channels %d must be a multiple of 8, <= 2048, with (c/8) | 256", who, c);
    return 0;
}

}  // namespace

extern "C" {

int tris_bn_apply_fwd(const void* y0, const float* stats0, const float* gamma0, const float* beta0, float* rm0, float* rv0,
                      float* save_mean0, float* save_invstd0, const void* y1, const float* stats1, const float* gamma1,
                      const float* beta1, float* rm1, float* rv1, float* save_mean1, float* save_invstd1,
                      const void* residual, void* out, int n, int h, int w, int c, int pool, int relu, int train,
                      float momentum, float eps, int stats_parts, int fold_half, float* save_scale0, float* save_shift0,
                      void* relu_bits, int unpair, tris_stream_t stream) {
    if (int e = check_c(c, "tris_bn_apply_fwd")) return e;
    if (pool != 1 && pool != 2) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: pool must be 1 or 2");
    if (pool == 2 && (y1 || residual || (h & 1) || (w & 1))) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: pooled form is single-branch, even h/w");
    if (train && (stats_parts < 1 || !stats0 || (y1 && !stats1)))
        return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: train mode needs the partial statistics rows");
    if (fold_half < 0 || (fold_half > 0 && 2 * fold_half != c)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: fold_half must be c/2");
    ApplyParams p{};
    p.b0 = {reinterpret_cast<const __nv_bfloat16*>(y0), stats0, gamma0, beta0, rm0, rv0, save_mean0, save_invstd0};
    p.b1 = {reinterpret_cast<const __nv_bfloat16*>(y1), stats1, gamma1, beta1, rm1, rv1, save_mean1, save_invstd1};
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.relu_bits = reinterpret_cast<unsigned char*>(relu_bits);
    if (relu_bits && (pool != 1 || !relu)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: relu_bits needs relu, pool 1");
    if (unpair && (pool != 2 || c % 16)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: unpair needs pool 2, c %% 16 == 0");
    p.n = n; p.h = h; p.w = w; p.c = c; p.pool = pool; p.relu = relu; p.train = train; p.unpair = unpair;
    p.count = static_cast<float>(static_cast<long>(n) * h * w);
    p.momentum = momentum; p.eps = eps;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    FinalizeParams f{};
    if (train) {
        f.parts[0] = stats0; f.parts[1] = stats1;
        f.rm[0] = rm0; f.rv[0] = rv0; f.mean[0] = save_mean0; f.invstd[0] = save_invstd0;
        f.rm[1] = rm1; f.rv[1] = rv1; f.mean[1] = save_mean1; f.invstd[1] = save_invstd1;
        f.gamma[0] = gamma0; f.beta[0] = beta0; f.gamma[1] = gamma1; f.beta[1] = beta1;
        f.scale[0] = save_scale0; f.shift[0] = save_shift0; f.scale[1] = nullptr; f.shift[1] = nullptr;
        f.nparts = stats_parts; f.c = c; f.fold_half = fold_half;
        f.count = p.count; f.momentum = momentum; f.eps = eps;
        TRIS_CUDA_OK(launch_dep(bn_fwd_finalize_kernel, dim3((c + 31) / 32, y1 ? 2 : 1), dim3(32 * kFinGroups), 0, s, f));
    }
    const long total = static_cast<long>(n) * (h / pool) * (w / pool) * (c / 8);
    if (train) {
        if (unpair) TRIS_CUDA_OK(launch_dep(bn_apply_fwd_unpair_kernel, dim3(grid_for(total)), dim3(kThreads), 4 * c * sizeof(float), s, p));
        else TRIS_CUDA_OK(launch_dep(bn_apply_fwd_kernel, dim3(grid_for(total)), dim3(kThreads), 4 * c * sizeof(float), s, p));
    } else {
        if (unpair) bn_apply_fwd_unpair_kernel<<<grid_for(total), kThreads, 4 * c * sizeof(float), s>>>(p);
        else bn_apply_fwd_kernel<<<grid_for(total), kThreads, 4 * c * sizeof(float), s>>>(p);
        TRIS_LAUNCH_OK("bn_apply_fwd_kernel");
    }
    return TRIS_OK;
}

/* Backward: [reduction kernel ->] finalize (in-order sum of the partial rows, += into dgamma / dbeta) -> apply.
 * ws: fp32 workspace [nparts][K][C] partial rows followed by [K][C] finalized sums (K = 3 with a second branch, else 2).
 * ext_parts > 0: the partial rows (K = 2: sum g | sum g (y - mean)) were already produced by the epilogue of the GEMM
 * that computed `dout` (tris_gemm stats_mode 1, dout = masked gradient g): the reduction kernel is skipped and `dout` is
 * used as g without a ReLU mask. */
int tris_bn_bwd(const void* dout, const void* out, const void* y0, const float* gamma0, const float* beta0,
                const float* save_mean0, const float* save_invstd0, float* dgamma0, float* dbeta0, void* dy0,
                const void* y1, const float* gamma1, const float* beta1, const float* save_mean1,
                const float* save_invstd1, float* dgamma1, float* dbeta1, void* dy1, void* g_out, int n, int h, int w,
                int c, int pool, int relu, int fold_half, float* ws, long ws_floats, int ext_parts,
                const void* relu_bits, int unpair, tris_stream_t stream) {
    if (int e = check_c(c, "tris_bn_bwd")) return e;
    if (pool != 1 && pool != 2) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: pool must be 1 or 2");
    if (!ws) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: workspace required");
    if (fold_half < 0 || (fold_half > 0 && 2 * fold_half != c)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: fold_half must be c/2");
    if (ext_parts > 0 && (y1 || pool != 1)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: external partial sums are single-branch, un-pooled");
    BwdParams p{};
    p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
    p.out = reinterpret_cast<const __nv_bfloat16*>(out);
    p.relu_bits = reinterpret_cast<const unsigned char*>(relu_bits);
    if (relu_bits && pool != 1) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: relu_bits is for the un-pooled form");
    p.b0 = {reinterpret_cast<const __nv_bfloat16*>(y0), nullptr, gamma0, beta0, nullptr, nullptr,
            const_cast<float*>(save_mean0), const_cast<float*>(save_invstd0)};
    p.b1 = {reinterpret_cast<const __nv_bfloat16*>(y1), nullptr, gamma1, beta1, nullptr, nullptr,
            const_cast<float*>(save_mean1), const_cast<float*>(save_invstd1)};
    p.dgamma0 = dgamma0; p.dbeta0 = dbeta0; p.dgamma1 = dgamma1; p.dbeta1 = dbeta1;
    p.dy0 = reinterpret_cast<__nv_bfloat16*>(dy0);
    p.dy1 = reinterpret_cast<__nv_bfloat16*>(dy1);
    p.g_out = reinterpret_cast<__nv_bfloat16*>(g_out);
    if (unpair && (pool != 2 || c % 16 || y1)) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: unpair needs pool 2, c %% 16 == 0, one branch");
    p.n = n; p.h = h; p.w = w; p.c = c; p.pool = pool; p.relu = ext_parts > 0 ? 0 : relu; p.unpair = unpair;
    p.count = static_cast<float>(static_cast<long>(n) * h * w);
    p.kred = y1 ? 3 : 2;
    p.fold_half = fold_half;
    const long total = static_cast<long>(n) * h * w * (c / 8);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rgrid = grid_for(total);
    if (rgrid > 4 * tris::sm_count()) rgrid = 4 * tris::sm_count();   // 4 resident CTAs/SM (64 registers/thread)
    const long per_row = static_cast<long>(p.kred) * c;
    if (ext_parts > 0) {
        rgrid = ext_parts;
    } else if ((rgrid + 1) * per_row > ws_floats) {
        rgrid = static_cast<int>(ws_floats / per_row) - 1;
    }
    if (rgrid < 1 || (rgrid + 1) * per_row > ws_floats) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: workspace of %ld floats is too small", ws_floats);
    p.parts = ws; p.nparts = rgrid; p.red = ws + rgrid * per_row;
    const size_t rsmem = (4 * c + kThreads * 8) * sizeof(float), asmem = 7 * c * sizeof(float);
    static bool attr = false;
    if (!attr) {
        TRIS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_apply_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4));
        TRIS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_apply_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4));
        TRIS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_apply_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4));
        attr = true;
    }
    if (ext_parts == 0) {
        if (y1 != nullptr) bn_bwd_reduce_kernel<true, false><<<rgrid, kThreads, rsmem, s>>>(p);
        else if (unpair) bn_bwd_reduce_kernel<false, true><<<rgrid, kThreads, rsmem, s>>>(p);
        else bn_bwd_reduce_kernel<false, false><<<rgrid, kThreads, rsmem, s>>>(p);
        TRIS_LAUNCH_OK("bn_bwd_reduce_kernel");
    }
    TRIS_CUDA_OK(launch_dep(bn_bwd_finalize_kernel, dim3((c + 31) / 32), dim3(32 * kFinGroupsBwd), 0, s, p));
    if (y1 != nullptr) TRIS_CUDA_OK(launch_dep((bn_bwd_apply_kernel<true, false>), dim3(grid_for(total)), dim3(kThreads), asmem, s, p));
    else if (unpair) TRIS_CUDA_OK(launch_dep((bn_bwd_apply_kernel<false, true>), dim3(grid_for(total)), dim3(kThreads), asmem, s, p));
    else TRIS_CUDA_OK(launch_dep((bn_bwd_apply_kernel<false, false>), dim3(grid_for(total)), dim3(kThreads), asmem, s, p));
    return TRIS_OK;
}

int tris_avgpool2_fwd(const void* x, void* out, int n, int h, int w, int c, tris_stream_t stream) {
    if (c % 8 || (h & 1) || (w & 1)) return tris::fail(TRIS_ERR_SHAPE, "tris_avgpool2_fwd: c%%8, even h/w");
    const long total = static_cast<long>(n) * (h / 2) * (w / 2) * (c / 8);
    avgpool2_fwd_kernel<<<grid_for(total), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), n, h, w, c);
    TRIS_LAUNCH_OK("avgpool2_fwd_kernel");
    return TRIS_OK;
}

int tris_avgpool2_bwd(const void* dout, const void* add, void* dx, int n, int h, int w, int c, tris_stream_t stream) {
    if (c % 8 || (h & 1) || (w & 1)) return tris::fail(TRIS_ERR_SHAPE, "tris_avgpool2_bwd: c%%8, even h/w");
    const long total = static_cast<long>(n) * h * w * (c / 8);
    avgpool2_bwd_kernel<<<grid_for(total), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(add),
        reinterpret_cast<__nv_bfloat16*>(dx), n, h, w, c);
    TRIS_LAUNCH_OK("avgpool2_bwd_kernel");
    return TRIS_OK;
}

}  // extern "C"
