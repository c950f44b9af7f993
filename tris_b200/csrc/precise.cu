// fp32 parity mode ("precision = fp32"): forward-only CUDA-core kernels that evaluate the Stage-1 network in full fp32
// (fp32 storage, fp32 FMA, fp64 BatchNorm sums), used to check the response maps against the reference's fp32 forward
// at 1e-3 (north-star tolerance) -- the training / throughput path is the bf16 tcgen05 one.  Everything here is a plain
// tiled kernel: the mode exists for fidelity, not speed (a single 320x320 image is ~22 GFLOP = a few ms).
//
// Restates the same reference lines as the bf16 kernels: CLIP/clip/model.py:10-55,212-279 (convs via im2col + SGEMM,
// BatchNorm train/eval, ReLU, AvgPool), :352-397,552-564 (LayerNorm, attention, QuickGELU MLP, embedding),
// model/attn.py:72-136 (InstanceNorm, softmaxes), model/model_stage1.py:61-78 (L2 norm, 0.1 mix).
#include <math.h>

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

// ---------------------------------------------------------------------------------------------- SGEMM
// C[M,N] (ldc) = act(alpha * A[M,K] . op(B) + bias[N]) + res[M,N];  A row-major (lda);  B is [N,K] row-major (b_kn = 0,
// nn.Linear layout) or [K,N] row-major (b_kn = 1).  64x64 tile, BK = 16, 256 threads, 4x4 micro-tile.
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                    const float* __restrict__ bias, const float* __restrict__ res, int M, int N, int K,
                                                    int lda, int ldb, int ldc, int b_kn, int act, float alpha, long sa, long sb, long sc) {
    __shared__ float As[SG_K][SG_T + 4], Bs[SG_K][SG_T + 4];
    A += blockIdx.z * sa; B += blockIdx.z * sb; C += blockIdx.z * sc;
    if (res != nullptr) res += blockIdx.z * sc;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += SG_K) {
        for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
            const int r = i / SG_K, kk = i % SG_K;          // K-contiguous operands: consecutive threads walk k
            const int m = m0 + r, k = k0 + kk;
            As[kk][r] = (m < M && k < K) ? A[static_cast<long>(m) * lda + k] : 0.f;
            if (!b_kn) {
                const int n = n0 + r;
                Bs[kk][r] = (n < N && k < K) ? B[static_cast<long>(n) * ldb + k] : 0.f;
            }
        }
        if (b_kn) {
            for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
                const int kk = i / SG_T, c = i % SG_T;
                const int n = n0 + c, k = k0 + kk;
                Bs[kk][c] = (n < N && k < K) ? B[static_cast<long>(k) * ldb + n] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] * alpha;
            if (bias != nullptr) v += bias[n];
            if (act == TRIS_ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == TRIS_ACT_QUICKGELU) v = v / (1.f + expf(-1.702f * v));
            if (res != nullptr) v += res[static_cast<long>(m) * ldc + n];
            C[static_cast<long>(m) * ldc + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------- conv helpers (NHWC fp32)
// x: NCHW (nchw = 1, the input image) or NHWC -> col [n*ho*wo, 9*C], k = (r*3+s)*C + c, pad 1, stride `st`.
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, float* __restrict__ col, int n, int h, int w, int C,
                                                        int st, int nchw) {
    const int ho = (h - 1) / st + 1, wo = (w - 1) / st + 1;
    const long total = static_cast<long>(n) * ho * wo * 9 * C;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int tap = static_cast<int>(t % 9);
        t /= 9;
        const int xo = static_cast<int>(t % wo);
        t /= wo;
        const int yo = static_cast<int>(t % ho);
        const long ni = t / ho;
        const int yi = yo * st + tap / 3 - 1, xi = xo * st + tap % 3 - 1;
        float v = 0.f;
        if (yi >= 0 && yi < h && xi >= 0 && xi < w)
            v = nchw ? x[((ni * C + c) * h + yi) * w + xi] : x[((ni * h + yi) * w + xi) * C + c];
        col[i] = v;
    }
}

// per-channel sum and sum of squares (fp64 accumulation) of x [rows, C]; stats[2C] pre-zeroed
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, double* __restrict__ stats, long rows, int C) {
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (long r = static_cast<long>(blockIdx.y) * 4 + part; r < rows; r += static_cast<long>(gridDim.y) * 4) {
        const double v = x[r * C + c];
        s += v;
        q += v * v;
    }
    atomicAdd(stats + c, s);
    atomicAdd(stats + C + c, q);
}

// out = [relu]( BN(y0) [+ BN(y1)] [+ res] ).  stats != NULL: batch statistics (train), else running mean / var (eval).
struct BnF32 {
    const float* y; const double* stats; const float* gamma; const float* beta; const float* rm; const float* rv;
};
__device__ __forceinline__ float bn_one(const BnF32& b, float v, int c, int C, double count, float eps) {
    float mean, var;
    if (b.stats != nullptr) {
        const double m = b.stats[c] / count;
        mean = static_cast<float>(m);
        var = static_cast<float>(fmax(b.stats[C + c] / count - m * m, 0.0));
    } else {
        mean = b.rm[c];
        var = b.rv[c];
    }
    return (v - mean) / sqrtf(var + eps) * b.gamma[c] + b.beta[c];
}
__global__ void __launch_bounds__(256) bn_f32_kernel(BnF32 b0, BnF32 b1, const float* __restrict__ res, float* __restrict__ out, long rows,
                                                     int C, int relu, float eps) {
    const long total = rows * C;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        float v = bn_one(b0, b0.y[i], c, C, static_cast<double>(rows), eps);
        if (b1.y != nullptr) v += bn_one(b1, b1.y[i], c, C, static_cast<double>(rows), eps);
        if (res != nullptr) v += res[i];
        if (relu) v = fmaxf(v, 0.f);
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256) avgpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int h, int w, int C) {
    const int ho = h / 2, wo = w / 2;
    const long total = static_cast<long>(n) * ho * wo * C;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int xo = static_cast<int>(t % wo);
        t /= wo;
        const int yo = static_cast<int>(t % ho);
        const long ni = t / ho;
        const float* p = x + ((ni * h + 2 * yo) * w + 2 * xo) * C + c;
        out[i] = 0.25f * (p[0] + p[C] + p[static_cast<long>(w) * C] + p[static_cast<long>(w) * C + C]);
    }
}

// ---------------------------------------------------------------------------------------------- transformer pieces
__global__ void embed_f32_kernel(const int* __restrict__ ids, const float* __restrict__ E, const float* __restrict__ P, float* __restrict__ x,
                                 int* __restrict__ eot, int n, int L, int D) {
    const int row = blockIdx.x, s = row / L, l = row % L;
    const int id = ids[row];
    for (int d = threadIdx.x; d < D; d += blockDim.x) x[static_cast<long>(row) * D + d] = E[static_cast<long>(id) * D + d] + P[l * D + d];
    if (l == 0 && threadIdx.x == 0 && eot != nullptr) {
        int best = 0, bv = ids[s * L];
        for (int j = 1; j < L; ++j) if (ids[s * L + j] > bv) { bv = ids[s * L + j]; best = j; }
        eot[s] = s * L + best;
    }
}

// one warp per row, two-pass fp32
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                            float* __restrict__ y, int rows, int D, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<long>(row) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; q += d * d; }
    const float inv = 1.f / sqrtf(warp_sum(q) / D + eps);
    for (int c = lane; c < D; c += 32) y[static_cast<long>(row) * D + c] = (xr[c] - mean) * inv * g[c] + b[c];
}

// softmax(q k^T / 8 [+causal]) v, qkv fp32 [n*L, 3D], head dim 64; one CTA per (sample, head), one warp per query row
__global__ void __launch_bounds__(128) attn_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int heads, int causal) {
    extern __shared__ float sm[];   // k [L][65], v [L][65], p [4][L]
    float* k = sm;
    float* v = sm + L * 65;
    float* p = v + L * 65;
    const int n = blockIdx.x / heads, h = blockIdx.x % heads, D = heads * 64;
    const long row0 = static_cast<long>(n) * L;
    for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {
        const int l = i >> 6, d = i & 63;
        k[l * 65 + d] = qkv[(row0 + l) * 3 * D + D + h * 64 + d];
        v[l * 65 + d] = qkv[(row0 + l) * 3 * D + 2 * D + h * 64 + d];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* pw = p + warp * L;
    for (int a = warp; a < L; a += 4) {
        const float* q = qkv + (row0 + a) * 3 * D + h * 64;
        const float q0 = q[lane] * 0.125f, q1 = q[lane + 32] * 0.125f;
        float mx = -INFINITY;
        for (int b = 0; b < L; ++b) {
            float s = warp_sum(q0 * k[b * 65 + lane] + q1 * k[b * 65 + lane + 32]);
            if (causal && b > a) s = -INFINITY;
            if (lane == 0) pw[b] = s;
            mx = fmaxf(mx, s);
        }
        __syncwarp();
        float den = 0.f;
        for (int b = lane; b < L; b += 32) { const float e = expf(pw[b] - mx); pw[b] = e; den += e; }
        den = 1.f / warp_sum(den);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int b = 0; b < L; ++b) { const float w = pw[b] * den; o0 = fmaf(w, v[b * 65 + lane], o0); o1 = fmaf(w, v[b * 65 + lane + 32], o1); }
        out[(row0 + a) * D + h * 64 + lane] = o0;
        out[(row0 + a) * D + h * 64 + lane + 32] = o1;
        __syncwarp();
    }
}

// tok[n,0,:] = cls + pos[0]; tok[n,1+p,:] = patch[n*(T-1)+p,:] + pos[1+p]   (CLIP/clip/model.py:436-441)
__global__ void vit_assemble_f32_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                        float* __restrict__ tok, int N, int T, int D) {
    const long total = static_cast<long>(N) * T * D;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i % D);
        const long r = i / D;
        const int t = static_cast<int>(r % T);
        const long n = r / T;
        tok[i] = (t == 0 ? cls[d] : patch[(n * (T - 1) + (t - 1)) * D + d]) + pos[t * D + d];
    }
}

__global__ void gather_rows_f32_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out, int rows, int D) {
    const long total = static_cast<long>(rows) * D;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x)
        out[i] = x[static_cast<long>(idx[i / D]) * D + i % D];
}

// ---------------------------------------------------------------------------------------------- head pieces
// y = x / ||x|| per row (one warp per row)
__global__ void __launch_bounds__(256) l2norm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = x[static_cast<long>(row) * D + c]; s += v * v; }
    const float inv = 1.f / sqrtf(warp_sum(s));
    for (int c = lane; c < D; c += 32) y[static_cast<long>(row) * D + c] = x[static_cast<long>(row) * D + c] * inv;
}

// out = mix_scale * act(IN(x) * gamma + beta) + mix_add ; x fp32 [B*P, ldx] (channel window of width C), stats over P rows
__global__ void __launch_bounds__(256) instnorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ mix_add,
                                                           float* __restrict__ out, int P, int C, float mix_scale, int relu, float eps) {
    __shared__ float red[4][64];
    const int c = blockIdx.x * 64 + threadIdx.x, b = blockIdx.y, pg = threadIdx.y;
    float s = 0.f;
    for (int p = pg; p < P; p += 4) s += x[(static_cast<long>(b) * P + p) * C + c];
    red[pg][threadIdx.x] = s;
    __syncthreads();
    const float mean = (red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]) / P;
    __syncthreads();
    float q = 0.f;
    for (int p = pg; p < P; p += 4) { const float d = x[(static_cast<long>(b) * P + p) * C + c] - mean; q += d * d; }
    red[pg][threadIdx.x] = q;
    __syncthreads();
    const float var = (red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]) / P;
    const float inv = 1.f / sqrtf(var + eps), g = gamma[c], be = beta[c];
    for (int p = pg; p < P; p += 4) {
        const long idx = (static_cast<long>(b) * P + p) * C + c;
        float o = (x[idx] - mean) * inv * g + be;
        if (relu) o = fmaxf(o, 0.f);
        o *= mix_scale;
        if (mix_add != nullptr) o += mix_add[idx];
        out[idx] = o;
    }
}

// row softmax of x [rows, n] (ld) * scale, in place capable
__global__ void __launch_bounds__(256) softmax_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int n, int ld, float scale) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = x + static_cast<long>(row) * ld;
    float mx = -INFINITY;
    for (int c = lane; c < n; c += 32) mx = fmaxf(mx, r[c] * scale);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < n; c += 32) s += expf(r[c] * scale - mx);
    s = 1.f / warp_sum(s);
    for (int c = lane; c < n; c += 32) y[static_cast<long>(row) * ld + c] = expf(r[c] * scale - mx) * s;
}

// out[b, i] = base[i] + a * x[b, i]
__global__ void bcast_mix_f32_kernel(const float* __restrict__ base, const float* __restrict__ x, float* __restrict__ out, long per, long total,
                                     float a) {
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x)
        out[i] = base[i % per] + a * x[i];
}

int grid1d(long n) {
    long g = (n + 255) / 256;
    const long cap = static_cast<long>(tris::sm_count()) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int tris_sgemm(const float* A, const float* B, float* C, const float* bias, const float* res, int M, int N, int K, int lda, int ldb, int ldc,
               int b_kn, int act, float alpha, int batch, long sa, long sb, long sc, tris_stream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || batch < 1) return tris::fail(TRIS_ERR_SHAPE, "tris_sgemm: M=%d N=%d K=%d batch=%d", M, N, K, batch);
    dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, batch);
    sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, bias, res, M, N, K, lda, ldb, ldc, b_kn, act, alpha == 0.f ? 1.f : alpha, sa, sb, sc);
    TRIS_LAUNCH_OK("sgemm_kernel");
    return TRIS_OK;
}

int tris_im2col3x3_f32(const float* x, float* col, int n, int h, int w, int c, int stride, int nchw, tris_stream_t stream) {
    const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
    im2col3x3_kernel<<<grid1d(static_cast<long>(n) * ho * wo * 9 * c), 256, 0, (cudaStream_t)stream>>>(x, col, n, h, w, c, stride, nchw);
    TRIS_LAUNCH_OK("im2col3x3_kernel");
    return TRIS_OK;
}

int tris_colstats_f32(const float* x, double* stats, long rows, int c, tris_stream_t stream) {
    TRIS_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * c * sizeof(double), (cudaStream_t)stream));
    int chunks = static_cast<int>((rows + 255) / 256);
    if (chunks > 256) chunks = 256;
    colstats_kernel<<<dim3((c + 63) / 64, chunks), 256, 0, (cudaStream_t)stream>>>(x, stats, rows, c);
    TRIS_LAUNCH_OK("colstats_kernel");
    return TRIS_OK;
}

int tris_bn_f32(const float* y0, const double* stats0, const float* gamma0, const float* beta0, const float* rm0, const float* rv0,
                const float* y1, const double* stats1, const float* gamma1, const float* beta1, const float* rm1, const float* rv1,
                const float* res, float* out, long rows, int c, int relu, float eps, tris_stream_t stream) {
    BnF32 b0{y0, stats0, gamma0, beta0, rm0, rv0}, b1{y1, stats1, gamma1, beta1, rm1, rv1};
    bn_f32_kernel<<<grid1d(rows * c), 256, 0, (cudaStream_t)stream>>>(b0, b1, res, out, rows, c, relu, eps);
    TRIS_LAUNCH_OK("bn_f32_kernel");
    return TRIS_OK;
}

int tris_avgpool2_f32(const float* x, float* out, int n, int h, int w, int c, tris_stream_t stream) {
    avgpool2_f32_kernel<<<grid1d(static_cast<long>(n) * (h / 2) * (w / 2) * c), 256, 0, (cudaStream_t)stream>>>(x, out, n, h, w, c);
    TRIS_LAUNCH_OK("avgpool2_f32_kernel");
    return TRIS_OK;
}

int tris_embed_f32(const int* ids, const float* E, const float* P, float* x, int* eot, int n, int L, int D, tris_stream_t stream) {
    embed_f32_kernel<<<n * L, 128, 0, (cudaStream_t)stream>>>(ids, E, P, x, eot, n, L, D);
    TRIS_LAUNCH_OK("embed_f32_kernel");
    return TRIS_OK;
}

int tris_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int rows, int D, float eps, tris_stream_t stream) {
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, rows, D, eps);
    TRIS_LAUNCH_OK("layernorm_f32_kernel");
    return TRIS_OK;
}

int tris_attn_f32(const float* qkv, float* out, int n, int L, int heads, int causal, tris_stream_t stream) {
    const size_t smem = (2 * static_cast<size_t>(L) * 65 + 4 * L) * sizeof(float);
    if (smem > 48 * 1024) return tris::fail(TRIS_ERR_SHAPE, "tris_attn_f32: L=%d too long", L);
    attn_f32_kernel<<<n * heads, 128, smem, (cudaStream_t)stream>>>(qkv, out, L, heads, causal);
    TRIS_LAUNCH_OK("attn_f32_kernel");
    return TRIS_OK;
}

int tris_vit_assemble_f32(const float* patch, const float* cls, const float* pos, float* tok, int n, int T, int D, tris_stream_t stream) {
    vit_assemble_f32_kernel<<<grid1d(static_cast<long>(n) * T * D), 256, 0, (cudaStream_t)stream>>>(patch, cls, pos, tok, n, T, D);
    TRIS_LAUNCH_OK("vit_assemble_f32_kernel");
    return TRIS_OK;
}

int tris_gather_rows_f32(const float* x, const int* idx, float* out, int rows, int D, tris_stream_t stream) {
    gather_rows_f32_kernel<<<grid1d(static_cast<long>(rows) * D), 256, 0, (cudaStream_t)stream>>>(x, idx, out, rows, D);
    TRIS_LAUNCH_OK("gather_rows_f32_kernel");
    return TRIS_OK;
}

int tris_l2norm_f32(const float* x, float* y, int rows, int D, tris_stream_t stream) {
    l2norm_f32_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, y, rows, D);
    TRIS_LAUNCH_OK("l2norm_f32_kernel");
    return TRIS_OK;
}

int tris_instnorm_f32(const float* x, const float* gamma, const float* beta, const float* mix_add, float* out, int batch, int P, int C,
                      float mix_scale, int relu, float eps, tris_stream_t stream) {
    if (C % 64) return tris::fail(TRIS_ERR_SHAPE, "tris_instnorm_f32: C=%d %% 64", C);
    instnorm_f32_kernel<<<dim3(C / 64, batch), dim3(64, 4), 0, (cudaStream_t)stream>>>(x, gamma, beta, mix_add, out, P, C, mix_scale, relu, eps);
    TRIS_LAUNCH_OK("instnorm_f32_kernel");
    return TRIS_OK;
}

int tris_softmax_f32(const float* x, float* y, int rows, int n, int ld, float scale, tris_stream_t stream) {
    softmax_f32_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, y, rows, n, ld, scale);
    TRIS_LAUNCH_OK("softmax_f32_kernel");
    return TRIS_OK;
}

int tris_bcast_mix_f32(const float* base, const float* x, float* out, long per, int B, float a, tris_stream_t stream) {
    bcast_mix_f32_kernel<<<grid1d(per * B), 256, 0, (cudaStream_t)stream>>>(base, x, out, per, per * B, a);
    TRIS_LAUNCH_OK("bcast_mix_f32_kernel");
    return TRIS_OK;
}

}  // extern "C"
