// Small HBM-bound kernels around the GEMMs: stem im2col (stride-2 3x3 on the fp32 NCHW image), fp32->bf16 weight
// shadow conversion and 3x3 weight (un)packing, row L2 normalisation, per-image InstanceNorm (+ReLU / residual mix)
// forward and backward, bf16 axpy.
//
// Replaces: first stem conv's data movement (CLIP/clip/model.py:212-217), the L2 normalisation of
// model/model_stage1.py:68-69, nn.InstanceNorm2d + ReLU of model/attn.py:72-104 and the 0.1-residual mix of
// model_stage1.py:73-74.
#include <cuda_bf16.h>

#include "common.h"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// img fp32 [N,3,H,W] -> col bf16 [N*Ho*Wo, 64]; k = (r*3+s)*3 + c for k < 27, zero for k >= 27. stride 2, pad 1.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col,
                                                          int N, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const long total = static_cast<long>(N) * Ho * Wo;
    for (long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long>(gridDim.x) * blockDim.x) {
        const int wo = static_cast<int>(pix % Wo);
        const long t = pix / Wo;
        const int ho = static_cast<int>(t % Ho);
        const long n = t / Ho;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hi = 2 * ho + r - 1;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int wi = 2 * wo + s - 1;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = __ldg(img + ((n * 3 + c) * H + hi) * W + wi);
                }
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(col + pix * 64);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            dst[g] = o;
        }
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int g = 4; g < 8; ++g) dst[g] = z;
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
    const long n4 = n >> 2;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src) + i);
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(dst) + 2 * i;
        o[0] = __floats2bfloat162_rn(f.x, f.y);
        o[1] = __floats2bfloat162_rn(f.z, f.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16(src[(n4 << 2) + threadIdx.x]);
}

// Image-pair-packed variant (stem, resnet.py): row = pixel of the image PAIR (2i, 2i+1); columns 0..26 = patch of image
// 2i, 32..58 = patch of image 2i+1, rest zero.  The 32-channel stem activations then carry two images in one 64-channel
// row (full 128-byte TMA rows with no padding) and the stem convs use block-diagonal weights.
__global__ void __launch_bounds__(256) stem_im2col_pair_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col,
                                                               int N2, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const long total = static_cast<long>(N2) * Ho * Wo * 2;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int half = static_cast<int>(i & 1);
        const long pix = i >> 1;
        const int wo = static_cast<int>(pix % Wo);
        const long t = pix / Wo;
        const int ho = static_cast<int>(t % Ho);
        const long n = (t / Ho) * 2 + half;
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hi = 2 * ho + r - 1;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int wi = 2 * wo + q - 1;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[(r * 3 + q) * 3 + c] = __ldg(img + ((n * 3 + c) * H + hi) * W + wi);
                }
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(col + pix * 64 + half * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            dst[g] = o;
        }
    }
}

// OIHW fp32 [co, ci, kh, kw] -> bf16 block-diagonal [reps*co, khw * reps*ci]: block (r, r) = the packed weight, rest zero.
__global__ void pack_conv_blockdiag_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int co, int ci, int khw, int reps) {
    const int CI = reps * ci;
    const long total = static_cast<long>(reps) * co * khw * CI;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int cc = static_cast<int>(i % CI);
        const long t = i / CI;
        const int tap = static_cast<int>(t % khw);
        const int oo = static_cast<int>(t / khw);
        const int ro = oo / co, o = oo % co, rc = cc / ci, c = cc % ci;
        out[i] = __float2bfloat16(ro == rc ? w[(static_cast<long>(o) * ci + c) * khw + tap] : 0.f);
    }
}
// gw[o, c, tap] += sum_r gp[(r*co + o), tap, r*ci + c]    (gp fp32 [reps*co, khw * reps*ci])
__global__ void unpack_conv_grad_blockdiag_kernel(const float* __restrict__ gp, float* __restrict__ gw, int co, int ci, int khw, int reps) {
    const long total = static_cast<long>(co) * ci * khw;
    const int CI = reps * ci;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % khw);
        const long t = i / khw;
        const int c = static_cast<int>(t % ci);
        const int o = static_cast<int>(t / ci);
        float s = 0.f;
        for (int r = 0; r < reps; ++r) s += gp[(static_cast<long>(r * co + o) * khw + tap) * CI + r * ci + c];
        gw[i] += s;
    }
}
// OIHW fp32 [co, ci, kh, kw] -> bf16 [co_pad, kh*kw*ci_pad] with k = (r*kw+s)*ci_pad + c (zero padding)
__global__ void pack_conv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int co, int ci, int khw, int co_pad,
                                 int ci_pad) {
    const long total = static_cast<long>(co_pad) * khw * ci_pad;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % ci_pad);
        const long t = i / ci_pad;
        const int tap = static_cast<int>(t % khw);
        const int o = static_cast<int>(t / khw);
        float v = 0.f;
        if (o < co && c < ci) v = w[(static_cast<long>(o) * ci + c) * khw + tap];
        out[i] = __float2bfloat16(v);
    }
}
// packed gradient fp32 [co_pad, khw*ci_pad] -> OIHW fp32 grad (+=)
__global__ void unpack_conv_grad_kernel(const float* __restrict__ gp, float* __restrict__ gw, int co, int ci, int khw, int ci_pad) {
    const long total = static_cast<long>(co) * ci * khw;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % khw);
        const long t = i / khw;
        const int c = static_cast<int>(t % ci);
        const int o = static_cast<int>(t / ci);
        gw[i] += gp[(static_cast<long>(o) * khw + tap) * ci_pad + c];
    }
}

// y = x / ||x||_2 per row (no eps, as the reference); inv_norm saved. one warp per row. x bf16 or fp32 in; bf16 out.
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ inv_norm, int rows, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const __nv_bfloat16* xr = x + static_cast<long>(row) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = __bfloat162float(xr[c]); s += v * v; }
    const float inv = rsqrtf(warp_sum(s));
    for (int c = lane; c < D; c += 32) y[static_cast<long>(row) * D + c] = __float2bfloat16(__bfloat162float(xr[c]) * inv);
    if (lane == 0) inv_norm[row] = inv;
}
// fp32 input variant: also emits the normalised rows in fp32 (input of the per-image centring below).
__global__ void __launch_bounds__(256) l2norm_fwd_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                             float* __restrict__ y32, float* __restrict__ inv_norm, int rows, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<long>(row) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = xr[c]; s += v * v; }
    const float inv = rsqrtf(warp_sum(s));
    for (int c = lane; c < D; c += 32) {
        const float v = xr[c] * inv;
        y[static_cast<long>(row) * D + c] = __float2bfloat16(v);
        if (y32 != nullptr) y32[static_cast<long>(row) * D + c] = v;
    }
    if (lane == 0) inv_norm[row] = inv;
}

// out (bf16) = x - mean over the P pixels of each image, per channel; x fp32 [B*P, C].  The centring is done in fp32
// BEFORE the bf16 rounding: the 1x1 convs that consume it are followed by an InstanceNorm (attn.py:72-86), which removes
// the pixel mean anyway, so feeding the centred tensor is exact -- and keeps the pixel-to-pixel variation of a nearly
// pixel-constant feature map at full bf16 relative precision.  block = (64 channels, 4 pixel groups), grid = (C/64, B).
__global__ void __launch_bounds__(256) center_pixels_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int P, int C) {
    __shared__ float red[4][64];
    const int c = blockIdx.x * 64 + threadIdx.x, b = blockIdx.y, pg = threadIdx.y;
    float s = 0.f;
    for (int p = pg; p < P; p += 4) s += x[(static_cast<long>(b) * P + p) * C + c];
    red[pg][threadIdx.x] = s;
    __syncthreads();
    const float mean = (red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]) / P;
    for (int p = pg; p < P; p += 4) {
        const long idx = (static_cast<long>(b) * P + p) * C + c;
        out[idx] = __float2bfloat16(x[idx] - mean);
    }
}

// dx = inv * (dy - y * <dy, y>)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                         const float* __restrict__ inv_norm, __nv_bfloat16* __restrict__ dx, int rows,
                                                         int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long base = static_cast<long>(row) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += __bfloat162float(dy[base + c]) * __bfloat162float(y[base + c]);
    s = warp_sum(s);
    const float inv = inv_norm[row];
    for (int c = lane; c < D; c += 32)
        dx[base + c] = __float2bfloat16(inv * (__bfloat162float(dy[base + c]) - __bfloat162float(y[base + c]) * s));
}

// InstanceNorm over the P pixels of each image, per channel.  x bf16 [B*P, ldx] (channel offset applied by caller),
// block = (64 channels, 4 pixel groups), grid = (C/64, B).   out = mix_scale * (IN(x)*gamma+beta [relu]) + mix_add
// per-thread pixel cache IN_MAXP: P <= 4 * IN_MAXP (32 -> 128 pixels = 320x320 inputs; 64 -> 256 pixels = up to 512x512)
template <int IN_MAXP>
__global__ void __launch_bounds__(256) instnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const __nv_bfloat16* __restrict__ mix_add,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out,
                                                           float* __restrict__ invstd_out, int P, int C, float mix_scale, int relu,
                                                           float eps) {
    __shared__ float red[2][4][64];
    const int c = blockIdx.x * 64 + threadIdx.x, b = blockIdx.y, pg = threadIdx.y;
    float v[IN_MAXP];
    float s = 0.f, q = 0.f;
    int cnt = 0;
    for (int p = pg; p < P; p += 4, ++cnt) {
        v[cnt] = __bfloat162float(x[(static_cast<long>(b) * P + p) * C + c]);
        s += v[cnt];
    }
    red[0][pg][threadIdx.x] = s;
    __syncthreads();
    const float mean = (red[0][0][threadIdx.x] + red[0][1][threadIdx.x] + red[0][2][threadIdx.x] + red[0][3][threadIdx.x]) / P;
    for (int i = 0; i < cnt; ++i) { const float d = v[i] - mean; q += d * d; }
    red[1][pg][threadIdx.x] = q;
    __syncthreads();
    const float var = (red[1][0][threadIdx.x] + red[1][1][threadIdx.x] + red[1][2][threadIdx.x] + red[1][3][threadIdx.x]) / P;
    const float invstd = rsqrtf(var + eps);
    const float g = gamma[c], be = beta[c];
    int i = 0;
    for (int p = pg; p < P; p += 4, ++i) {
        float o = (v[i] - mean) * invstd * g + be;
        if (relu) o = fmaxf(o, 0.f);
        o *= mix_scale;
        const long idx = (static_cast<long>(b) * P + p) * C + c;
        if (mix_add != nullptr) o += __bfloat162float(mix_add[idx]);
        out[idx] = __float2bfloat16(o);
    }
    if (pg == 0 && mean_out != nullptr) { mean_out[b * C + c] = mean; invstd_out[b * C + c] = invstd; }
}

// backward of out = mix_scale * act(IN(x)*gamma+beta): dx, and per-image partial rows of dgamma/dbeta in ws[2][B][C] (plain
// stores; the caller's queue adds the B rows in order).  relu mask recomputed from x.
template <int IN_MAXP>
__global__ void __launch_bounds__(256) instnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean_in, const float* __restrict__ invstd_in,
                                                           __nv_bfloat16* __restrict__ dx, float* __restrict__ ws,
                                                           int P, int C, float mix_scale, int relu) {
    __shared__ float red[2][4][64];
    const int c = blockIdx.x * 64 + threadIdx.x, b = blockIdx.y, pg = threadIdx.y;
    const float mean = mean_in[b * C + c], invstd = invstd_in[b * C + c], g = gamma[c], be = beta[c];
    float gz[IN_MAXP], xh[IN_MAXP];
    float s1 = 0.f, s2 = 0.f;
    int cnt = 0;
    for (int p = pg; p < P; p += 4, ++cnt) {
        const long idx = (static_cast<long>(b) * P + p) * C + c;
        xh[cnt] = (__bfloat162float(x[idx]) - mean) * invstd;
        float d = __bfloat162float(dout[idx]) * mix_scale;
        if (relu && (xh[cnt] * g + be) <= 0.f) d = 0.f;
        gz[cnt] = d;
        s1 += d;
        s2 += d * xh[cnt];
    }
    red[0][pg][threadIdx.x] = s1;
    red[1][pg][threadIdx.x] = s2;
    __syncthreads();
    s1 = red[0][0][threadIdx.x] + red[0][1][threadIdx.x] + red[0][2][threadIdx.x] + red[0][3][threadIdx.x];
    s2 = red[1][0][threadIdx.x] + red[1][1][threadIdx.x] + red[1][2][threadIdx.x] + red[1][3][threadIdx.x];
    if (pg == 0 && ws != nullptr) {
        ws[(static_cast<long>(gridDim.y) + b) * C + c] = s1;      // plane 1: dbeta
        ws[static_cast<long>(b) * C + c] = s2;                     // plane 0: dgamma
    }
    const float a = g * invstd, m1 = s1 / P, m2 = s2 / P;
    int i = 0;
    for (int p = pg; p < P; p += 4, ++i) {
        const long idx = (static_cast<long>(b) * P + p) * C + c;
        dx[idx] = __float2bfloat16(a * (gz[i] - m1 - xh[i] * m2));
    }
}

// y = a*x + b*y (bf16), 8 elements per thread
__global__ void __launch_bounds__(256) axpby_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float a, float b,
                                                    long n8) {
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
        uint4 xv = __ldg(reinterpret_cast<const uint4*>(x) + i);
        uint4 yv = reinterpret_cast<uint4*>(y)[i];
        __nv_bfloat162* xh = reinterpret_cast<__nv_bfloat162*>(&xv);
        __nv_bfloat162* yh = reinterpret_cast<__nv_bfloat162*>(&yv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xf = __bfloat1622float2(xh[j]), yf = __bfloat1622float2(yh[j]);
            yh[j] = __floats2bfloat162_rn(a * xf.x + b * yf.x, a * xf.y + b * yf.y);
        }
        reinterpret_cast<uint4*>(y)[i] = yv;
    }
}

int grid1d(long n, int cap_mult = 8) {
    long b = (n + 255) / 256;
    const long cap = static_cast<long>(tris::sm_count()) * cap_mult;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int tris_stem_im2col(const float* img, void* col, int n, int h, int w, tris_stream_t stream) {
    if ((h & 1) || (w & 1)) return tris::fail(TRIS_ERR_SHAPE, "tris_stem_im2col: even h/w required");
    stem_im2col_kernel<<<grid1d(static_cast<long>(n) * (h / 2) * (w / 2)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        img, reinterpret_cast<__nv_bfloat16*>(col), n, h, w);
    TRIS_LAUNCH_OK("stem_im2col_kernel");
    return TRIS_OK;
}

int tris_stem_im2col_pair(const float* img, void* col, int n, int h, int w, tris_stream_t stream) {
    if ((h & 1) || (w & 1) || (n & 1)) return tris::fail(TRIS_ERR_SHAPE, "tris_stem_im2col_pair: even n/h/w required");
    stem_im2col_pair_kernel<<<grid1d(static_cast<long>(n) * (h / 2) * (w / 2)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        img, reinterpret_cast<__nv_bfloat16*>(col), n / 2, h, w);
    TRIS_LAUNCH_OK("stem_im2col_pair_kernel");
    return TRIS_OK;
}

int tris_pack_conv_blockdiag(const float* w, void* out, int co, int ci, int khw, int reps, tris_stream_t stream) {
    pack_conv_blockdiag_kernel<<<grid1d(static_cast<long>(reps) * co * khw * reps * ci), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), co, ci, khw, reps);
    TRIS_LAUNCH_OK("pack_conv_blockdiag_kernel");
    return TRIS_OK;
}

int tris_unpack_conv_grad_blockdiag(const float* gp, float* gw, int co, int ci, int khw, int reps, tris_stream_t stream) {
    unpack_conv_grad_blockdiag_kernel<<<grid1d(static_cast<long>(co) * ci * khw), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        gp, gw, co, ci, khw, reps);
    TRIS_LAUNCH_OK("unpack_conv_grad_blockdiag_kernel");
    return TRIS_OK;
}

int tris_f32_to_bf16(const float* src, void* dst, long n, tris_stream_t stream) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7))
        return tris::fail(TRIS_ERR_ALIGN, "tris_f32_to_bf16: unaligned");
    f32_to_bf16_kernel<<<grid1d(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    TRIS_LAUNCH_OK("f32_to_bf16_kernel");
    return TRIS_OK;
}

int tris_pack_conv(const float* w, void* out, int co, int ci, int khw, int co_pad, int ci_pad, tris_stream_t stream) {
    pack_conv_kernel<<<grid1d(static_cast<long>(co_pad) * khw * ci_pad), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), co, ci, khw, co_pad, ci_pad);
    TRIS_LAUNCH_OK("pack_conv_kernel");
    return TRIS_OK;
}

int tris_unpack_conv_grad(const float* gp, float* gw, int co, int ci, int khw, int ci_pad, tris_stream_t stream) {
    unpack_conv_grad_kernel<<<grid1d(static_cast<long>(co) * ci * khw), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gp, gw, co, ci,
                                                                                                                    khw, ci_pad);
    TRIS_LAUNCH_OK("unpack_conv_grad_kernel");
    return TRIS_OK;
}

int tris_l2norm_fwd(const void* x, void* y, float* inv_norm, int rows, int D, tris_stream_t stream) {
    l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), inv_norm, rows, D);
    TRIS_LAUNCH_OK("l2norm_fwd_kernel");
    return TRIS_OK;
}

int tris_l2norm_fwd_f32(const float* x, void* y, float* y32, float* inv_norm, int rows, int D, tris_stream_t stream) {
    l2norm_fwd_f32_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), y32,
                                                                                              inv_norm, rows, D);
    TRIS_LAUNCH_OK("l2norm_fwd_f32_kernel");
    return TRIS_OK;
}

int tris_center_pixels(const float* x, void* out, int B, int P, int C, tris_stream_t stream) {
    if (C % 64) return tris::fail(TRIS_ERR_SHAPE, "center_pixels: C=%d %% 64", C);
    center_pixels_kernel<<<dim3(C / 64, B), dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(out), P, C);
    TRIS_LAUNCH_OK("center_pixels_kernel");
    return TRIS_OK;
}

int tris_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int rows, int D, tris_stream_t stream) {
    l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(y), inv_norm,
        reinterpret_cast<__nv_bfloat16*>(dx), rows, D);
    TRIS_LAUNCH_OK("l2norm_bwd_kernel");
    return TRIS_OK;
}

int tris_instnorm_fwd(const void* x, const float* gamma, const float* beta, const void* mix_add, void* out, float* mean,
                      float* invstd, int batch, int P, int C, float mix_scale, int relu, float eps, tris_stream_t stream) {
    if (C % 64 || P > 256) return tris::fail(TRIS_ERR_SHAPE, "tris_instnorm_fwd: C%%64, P<=256 (got C=%d P=%d)", C, P);
    auto fn = P <= 128 ? instnorm_fwd_kernel<32> : instnorm_fwd_kernel<64>;
    fn<<<dim3(C / 64, batch), dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<const __nv_bfloat16*>(mix_add),
        reinterpret_cast<__nv_bfloat16*>(out), mean, invstd, P, C, mix_scale, relu, eps);
    TRIS_LAUNCH_OK("instnorm_fwd_kernel");
    return TRIS_OK;
}

int tris_instnorm_bwd(const void* dout, const void* x, const float* gamma, const float* beta, const float* mean,
                      const float* invstd, void* dx, float* ws, int batch, int P, int C, float mix_scale,
                      int relu, tris_stream_t stream) {
    if (C % 64 || P > 256) return tris::fail(TRIS_ERR_SHAPE, "tris_instnorm_bwd: C%%64, P<=256 (got C=%d P=%d)", C, P);
    auto fn = P <= 128 ? instnorm_bwd_kernel<32> : instnorm_bwd_kernel<64>;
    fn<<<dim3(C / 64, batch), dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, mean, invstd,
        reinterpret_cast<__nv_bfloat16*>(dx), ws, P, C, mix_scale, relu);
    TRIS_LAUNCH_OK("instnorm_bwd_kernel");
    return TRIS_OK;
}

int tris_axpby(const void* x, void* y, float a, float b, long n, tris_stream_t stream) {
    if (n % 8) return tris::fail(TRIS_ERR_SHAPE, "tris_axpby: n %% 8");
    axpby_kernel<<<grid1d(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), a, b, n / 8);
    TRIS_LAUNCH_OK("axpby_kernel");
    return TRIS_OK;
}

}  // extern "C"
