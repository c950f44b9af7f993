// BatchNorm2d (train/eval) + ReLU + residual add + 2x2 average pool on NHWC bf16 activations: forward apply,
// backward reduction and backward apply.  HBM-bound streaming kernels: 128-bit loads/stores (8 channels per
// thread), per-channel scale/shift staged in shared memory, fp32 math, warp/block reductions + fp32 atomics.
//
// Replaces nn.BatchNorm2d / ReLU / AvgPool2d / "out += identity" of CLIP/clip/model.py:18-55, :212-232, :255-260.
// Batch statistics (sum, sum of squares) arrive from the conv GEMM epilogue (gemm_sm100.cu).
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.h"

namespace {

constexpr int kThreads = 256;

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
    int n, h, w, c, pool, relu, train;
    float count, momentum, eps;
};

__device__ __forceinline__ void bn_prologue(const BnBranch& b, const ApplyParams& p, float* s_scale, float* s_shift) {
    for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
        float mean, var;
        if (p.train) {
            mean = b.stats[c] / p.count;
            var = fmaxf(b.stats[p.c + c] / p.count - mean * mean, 0.f);
        } else {
            mean = b.running_mean[c];
            var = b.running_var[c];
        }
        const float invstd = rsqrtf(var + p.eps);
        const float sc = b.gamma[c] * invstd;
        s_scale[c] = sc;
        s_shift[c] = b.beta[c] - mean * sc;
        if (p.train && blockIdx.x == 0) {
            b.save_mean[c] = mean;
            b.save_invstd[c] = invstd;
            if (b.running_mean != nullptr) {
                const float unbiased = var * (p.count / fmaxf(p.count - 1.f, 1.f));
                b.running_mean[c] = (1.f - p.momentum) * b.running_mean[c] + p.momentum * mean;
                b.running_var[c] = (1.f - p.momentum) * b.running_var[c] + p.momentum * unbiased;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) bn_apply_fwd_kernel(const ApplyParams p) {
    extern __shared__ float sm[];
    float* sc0 = sm;
    float* sh0 = sm + p.c;
    float* sc1 = sm + 2 * p.c;
    float* sh1 = sm + 3 * p.c;
    const bool dual = p.b1.y != nullptr;
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
        st8(p.out + orow * p.c + c0, acc);
    }
}

// ------------------------------------------------------------------------------------------------ backward
struct BwdParams {
    const __nv_bfloat16* dout;   // [M_out, C]
    const __nv_bfloat16* out;    // [M_out, C] forward output (relu mask source) or nullptr -> recompute from y0
    BnBranch b0, b1;             // y, gamma, beta, save_mean, save_invstd are inputs here
    float* dgamma0; float* dbeta0; float* dgamma1; float* dbeta1;   // [C] fp32, atomically accumulated
    __nv_bfloat16* dy0; __nv_bfloat16* dy1;   // [M_in, C]
    __nv_bfloat16* g_out;        // optional [M_out, C]: masked upstream gradient (identity branch)
    int n, h, w, c, pool, relu;
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
__device__ __forceinline__ Vec8 masked_grad(const BwdParams& p, long orow, int c0, const Vec8& y0, const float* s_sc, const float* s_sh) {
    Vec8 g = ld8(p.dout + orow * p.c + c0);
    if (p.pool == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g.v[i] *= 0.25f;
    }
    if (p.relu) {
        if (p.out != nullptr) {
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
template <bool DUAL>
__global__ void __launch_bounds__(kThreads, 4) bn_bwd_reduce_kernel(const BwdParams p) {
    extern __shared__ float sm[];
    float* s_sc = sm;
    float* s_sh = sm + p.c;
    float* s_mu0 = sm + 2 * p.c;
    float* s_mu1 = sm + 3 * p.c;
    float* red = sm + 4 * p.c;   // [kThreads * 8]
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
        const Vec8 g = masked_grad(p, orow, c0, y0, s_sc, s_sh);
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
    // block reduction over the kThreads/vecs threads that share a channel group
    const int per = kThreads / vecs;   // threads per channel group (>= 1)
    auto block_sum = [&](float (&x)[8], float* dst, const float* scale) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = x[i];
        __syncthreads();
        for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
            const int g = c >> 3, i = c & 7;
            float s = 0.f;
            for (int k = 0; k < per; ++k) s += red[(g + k * vecs) * 8 + i];
            atomicAdd(dst + c, scale != nullptr ? s * scale[c] : s);
        }
    };
    block_sum(db, p.dbeta0, nullptr);
    block_sum(dg0, p.dgamma0, p.b0.save_invstd);
    if (DUAL) {
        block_sum(db, p.dbeta1, nullptr);
        block_sum(dg1, p.dgamma1, p.b1.save_invstd);
    }
}

// dy = a g + b y + c per channel with  a = gamma*invstd,  b = -a*invstd*dgamma/N,  c = a*(invstd*dgamma/N*mean - dbeta/N).
// smem: sc0, sh0, b0, c0, (dual) a1, b1, c1  -- each [C].
template <bool DUAL>
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
    for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
        const float invstd = p.b0.save_invstd[c], mean = p.b0.save_mean[c];
        const float sc = p.b0.gamma[c] * invstd;
        const float k = p.dgamma0[c] * inv_count * invstd, m = p.dbeta0[c] * inv_count;
        s_sc[c] = sc; s_sh[c] = p.b0.beta[c] - mean * sc;
        s_b0[c] = -sc * k; s_c0[c] = sc * (k * mean - m);
        if (DUAL) {
            const float is1 = p.b1.save_invstd[c], mu1 = p.b1.save_mean[c];
            const float a1 = p.b1.gamma[c] * is1;
            const float k1 = p.dgamma1[c] * inv_count * is1, m1 = p.dbeta1[c] * inv_count;
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
        const Vec8 g = masked_grad(p, orow, c0, y0, s_sc, s_sh);
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
                      float momentum, float eps, tris_stream_t stream) {
    if (int e = check_c(c, "tris_bn_apply_fwd")) return e;
    if (pool != 1 && pool != 2) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: pool must be 1 or 2");
    if (pool == 2 && (y1 || residual || (h & 1) || (w & 1))) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_apply_fwd: pooled form is single-branch, even h/w");
    ApplyParams p{};
    p.b0 = {reinterpret_cast<const __nv_bfloat16*>(y0), stats0, gamma0, beta0, rm0, rv0, save_mean0, save_invstd0};
    p.b1 = {reinterpret_cast<const __nv_bfloat16*>(y1), stats1, gamma1, beta1, rm1, rv1, save_mean1, save_invstd1};
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.n = n; p.h = h; p.w = w; p.c = c; p.pool = pool; p.relu = relu; p.train = train;
    p.count = static_cast<float>(static_cast<long>(n) * h * w);
    p.momentum = momentum; p.eps = eps;
    const long total = static_cast<long>(n) * (h / pool) * (w / pool) * (c / 8);
    bn_apply_fwd_kernel<<<grid_for(total), kThreads, 4 * c * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(p);
    TRIS_LAUNCH_OK("bn_apply_fwd_kernel");
    return TRIS_OK;
}

/* Launches the reduction (dgamma/dbeta, atomically accumulated into pre-zeroed buffers) and the apply kernel. */
int tris_bn_bwd(const void* dout, const void* out, const void* y0, const float* gamma0, const float* beta0,
                const float* save_mean0, const float* save_invstd0, float* dgamma0, float* dbeta0, void* dy0,
                const void* y1, const float* gamma1, const float* beta1, const float* save_mean1,
                const float* save_invstd1, float* dgamma1, float* dbeta1, void* dy1, void* g_out, int n, int h, int w,
                int c, int pool, int relu, int fold_half, tris_stream_t stream) {
    if (int e = check_c(c, "tris_bn_bwd")) return e;
    if (pool != 1 && pool != 2) return tris::fail(TRIS_ERR_SHAPE, "tris_bn_bwd: pool must be 1 or 2");
    BwdParams p{};
    p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
    p.out = reinterpret_cast<const __nv_bfloat16*>(out);
    p.b0 = {reinterpret_cast<const __nv_bfloat16*>(y0), nullptr, gamma0, beta0, nullptr, nullptr,
            const_cast<float*>(save_mean0), const_cast<float*>(save_invstd0)};
    p.b1 = {reinterpret_cast<const __nv_bfloat16*>(y1), nullptr, gamma1, beta1, nullptr, nullptr,
            const_cast<float*>(save_mean1), const_cast<float*>(save_invstd1)};
    p.dgamma0 = dgamma0; p.dbeta0 = dbeta0; p.dgamma1 = dgamma1; p.dbeta1 = dbeta1;
    p.dy0 = reinterpret_cast<__nv_bfloat16*>(dy0);
    p.dy1 = reinterpret_cast<__nv_bfloat16*>(dy1);
    p.g_out = reinterpret_cast<__nv_bfloat16*>(g_out);
    p.n = n; p.h = h; p.w = w; p.c = c; p.pool = pool; p.relu = relu;
    p.count = static_cast<float>(static_cast<long>(n) * h * w);
    const long total = static_cast<long>(n) * h * w * (c / 8);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // cap the reduction grid: each block ends with 2-4 * C atomics
    int rgrid = grid_for(total);
    if (rgrid > 4 * tris::sm_count()) rgrid = 4 * tris::sm_count();   // 4 resident CTAs/SM (64 registers/thread)
    const size_t rsmem = (4 * c + kThreads * 8) * sizeof(float), asmem = 7 * c * sizeof(float);
    static bool attr = false;
    if (!attr) {
        TRIS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4));
        TRIS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4));
        attr = true;
    }
    if (y1 != nullptr) bn_bwd_reduce_kernel<true><<<rgrid, kThreads, rsmem, s>>>(p);
    else bn_bwd_reduce_kernel<false><<<rgrid, kThreads, rsmem, s>>>(p);
    TRIS_LAUNCH_OK("bn_bwd_reduce_kernel");
    if (fold_half > 0) {   // pair-packed stem: channels c and c + fold_half are one BatchNorm channel
        if (int e = tris_fold_pairs(dgamma0, dbeta0, nullptr, fold_half, stream)) return e;
    }
    if (y1 != nullptr) bn_bwd_apply_kernel<true><<<grid_for(total), kThreads, asmem, s>>>(p);
    else bn_bwd_apply_kernel<false><<<grid_for(total), kThreads, asmem, s>>>(p);
    TRIS_LAUNCH_OK("bn_bwd_apply_kernel");
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
