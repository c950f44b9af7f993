// Stage-1 head kernels around the cross-modal GEMMs (all small, HBM / latency bound; warp-shuffle reductions):
//   * the two softmaxes of the bilateral attention (model/attn.py:122,125) and their backward,
//   * the 0.1-residual mix of the per-image text features (model/model_stage1.py:74) and the sum over images,
//   * the response head: background class, 49-way softmax, mean/max/focal classification scores, diagonal map
//     (model/model_stage1.py:80-114, focal_loss :122-123) forward and backward,
//   * bilinear x32 up-sampling (align_corners=False, model/utils.py:5-10) + ReLU / sigmoid, forward and backward,
//   * mask-and-resize: fg = bilinear_ac(sigmoid map -> 224) * bilinear_ac(img -> 224) written straight as ViT patches
//     (train_stage1.py:327-339) forward and backward, plus the plain NCHW -> patch layout change,
//   * the three loss terms and their gradients (train_stage1.py:263-284, 340-364).
#include <cuda_bf16.h>
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

// ------------------------------------------------------------------------------------------------ attention softmaxes
// S1, S2T: f32 [B, P, Tp] (Tp = T rounded up to 8; columns >= T are ignored).
//   PA [b,p,:]  = softmax_t(S1[b,p,:]  * scale)      (text axis, attn.py:122)
//   PTt[b,:,t]  = softmax_p(S2T[b,:,t] * scale)      (pixel axis, attn.py:125; stored transposed)
//   PAc[b,p,t]  = PA[b,p,t] - mean_p' PA[b,p',t]     (computed in fp32 before the bf16 rounding)
// PAc replaces PA as the operand of PA.Vt: the InstanceNorm that follows v_output (attn.py:102-105) removes the pixel
// mean of every channel, so IN(W_o(PA Vt)+b) == IN(W_o(PAc Vt)) exactly, but the centred operand keeps the (small)
// pixel-to-pixel variation at full bf16 relative precision instead of burying it under the rounding of the mean.
// outputs bf16 [B, P, Tp], padded columns zero.  One CTA per image.
__global__ void __launch_bounds__(256) xattn_softmax_fwd_kernel(const float* __restrict__ S1, const float* __restrict__ S2T,
                                                                __nv_bfloat16* __restrict__ PA, __nv_bfloat16* __restrict__ PAc,
                                                                __nv_bfloat16* __restrict__ PTt, int P, int T, int Tp, float scale) {
    extern __shared__ float sm[];   // a [P*Tp]: row softmax; c [P*Tp]: S2T then the column softmax; colmean [Tp]
    float* a = sm;
    float* c = sm + P * Tp;
    float* colmean = c + P * Tp;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const long base = static_cast<long>(b) * P * Tp;
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) c[i] = S2T[base + i] * scale;
    for (int p = warp; p < P; p += nw) {
        const float* r = S1 + base + static_cast<long>(p) * Tp;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, r[t] * scale);
        mx = warp_max(mx);
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += __expf(r[t] * scale - mx);
        s = 1.f / warp_sum(s);
        for (int t = lane; t < Tp; t += 32) a[p * Tp + t] = t < T ? __expf(r[t] * scale - mx) * s : 0.f;
    }
    __syncthreads();
    for (int t = warp; t < Tp; t += nw) {
        float cs = 0.f;
        for (int p = lane; p < P; p += 32) cs += a[p * Tp + t];
        cs = warp_sum(cs);
        if (lane == 0) colmean[t] = cs / P;
        if (t >= T) continue;
        float mx = -INFINITY;
        for (int p = lane; p < P; p += 32) mx = fmaxf(mx, c[p * Tp + t]);
        mx = warp_max(mx);
        float s = 0.f;
        for (int p = lane; p < P; p += 32) s += __expf(c[p * Tp + t] - mx);
        s = 1.f / warp_sum(s);
        for (int p = lane; p < P; p += 32) c[p * Tp + t] = __expf(c[p * Tp + t] - mx) * s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) {
        const int t = i % Tp;
        PA[base + i] = __float2bfloat16(a[i]);
        PAc[base + i] = __float2bfloat16(a[i] - colmean[t]);
        PTt[base + i] = __float2bfloat16(t < T ? c[i] : 0.f);
    }
}

// dS = scale * P o (dP - sum(dP o P)) along the softmax axis; dPAc/dPTt f32 [B,P,Tp] in, bf16 [B,P,Tp] out.
// dPAc is the gradient w.r.t. the centred PAc: dPA = dPAc - mean_p dPAc (centring is a symmetric projection).
__global__ void __launch_bounds__(256) xattn_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ PA, const float* __restrict__ dPAc,
                                                                const __nv_bfloat16* __restrict__ PTt, const float* __restrict__ dPTt,
                                                                __nv_bfloat16* __restrict__ dS1, __nv_bfloat16* __restrict__ dS2T,
                                                                int P, int T, int Tp, float scale) {
    extern __shared__ float sm[];   // dots [T], colmean [T]
    float* dots = sm;
    float* colmean = sm + T;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const long base = static_cast<long>(b) * P * Tp;
    for (int t = warp; t < T; t += nw) {
        float d = 0.f, m = 0.f;
        for (int p = lane; p < P; p += 32) {
            d += dPTt[base + p * Tp + t] * __bfloat162float(PTt[base + p * Tp + t]);
            m += dPAc[base + p * Tp + t];
        }
        d = warp_sum(d);
        m = warp_sum(m);
        if (lane == 0) { dots[t] = d; colmean[t] = m / P; }
    }
    __syncthreads();
    for (int p = warp; p < P; p += nw) {
        const long o = base + static_cast<long>(p) * Tp;
        float d = 0.f;
        for (int t = lane; t < T; t += 32) d += (dPAc[o + t] - colmean[t]) * __bfloat162float(PA[o + t]);
        d = warp_sum(d);
        for (int t = lane; t < Tp; t += 32)
            dS1[o + t] = __float2bfloat16(t < T ? scale * __bfloat162float(PA[o + t]) * (dPAc[o + t] - colmean[t] - d) : 0.f);
    }
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) {
        const int t = i % Tp;
        dS2T[base + i] = __float2bfloat16(t < T ? scale * __bfloat162float(PTt[base + i]) * (dPTt[base + i] - dots[t]) : 0.f);
    }
}

// out[b, i] = base[i] + a * x[b, i]   (bf16; i over per = T*C elements; 8 per thread)
__global__ void __launch_bounds__(256) bcast_mix_kernel(const __nv_bfloat16* __restrict__ basev, const __nv_bfloat16* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ out, long per8, long total8, float a) {
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const uint4 xv = reinterpret_cast<const uint4*>(x)[i];
        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(basev) + (i % per8));
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xv);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bv);
        uint4 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xf = __bfloat1622float2(x2[j]), bf = __bfloat1622float2(b2[j]);
            o2[j] = __floats2bfloat162_rn(bf.x + a * xf.x, bf.y + a * xf.y);
        }
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// out[i] (bf16) = a * sum_b x[b, i]  (+ add[i])
__global__ void __launch_bounds__(256) batch_sum_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ add,
                                                        __nv_bfloat16* __restrict__ out, long per, int B, float a) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __bfloat162float(x[static_cast<long>(b) * per + i]);
    s *= a;
    if (add != nullptr) s += __bfloat162float(add[i]);
    out[i] = __float2bfloat16(s);
}

// dst (bf16) = g (f32) masked by y > 0   (ReLU backward of the text projections, attn.py:87-97)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                                        __nv_bfloat16* __restrict__ dst, long n) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(__bfloat162float(y[i]) > 0.f ? g[i] : 0.f);
}

// ------------------------------------------------------------------------------------------------ response head (K8)
// R f32 [B, P, Tp] = <v', l'> (unscaled); z = exp(logit_scale) * R.  One CTA per image.
//   cls_out[b,j] = mean_p z + max_p z + (1-m)^fp * log(fl + m),  m = mean_p softmax_{bg=1, z[p,:]}[j]
//   cls_fg[b]    = m[j = b]          maps[b,p] = z[p, b]
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ R, const float* __restrict__ logit_scale,
                                                       float* __restrict__ cls_out, float* __restrict__ cls_fg,
                                                       float* __restrict__ maps, float* __restrict__ mbar_out,
                                                       int* __restrict__ argmax_out, float* __restrict__ es_out, int P, int T, int Tp,
                                                       float focal_p, float focal_l, int train) {
    extern __shared__ float sm[];   // z [P*Tp], rmax [P], rinv [P]
    float* z = sm;
    float* rmax = sm + P * Tp;
    float* rinv = rmax + P;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float es = __expf(*logit_scale);
    const long base = static_cast<long>(b) * P * Tp;
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) z[i] = R[base + i] * es;
    if (b == 0 && threadIdx.x == 0 && es_out != nullptr) *es_out = es;
    __syncthreads();
    if (b < T)
        for (int p = threadIdx.x; p < P; p += blockDim.x) maps[static_cast<long>(b) * P + p] = z[p * Tp + b];
    if (!train) return;
    for (int p = warp; p < P; p += nw) {
        float mx = 1.f;   // background logit
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, z[p * Tp + t]);
        mx = warp_max(mx);
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += __expf(z[p * Tp + t] - mx);
        s = warp_sum(s) + __expf(1.f - mx);
        if (lane == 0) { rmax[p] = mx; rinv[p] = 1.f / s; }
    }
    __syncthreads();
    for (int t = warp; t < T; t += nw) {
        float sz = 0.f, mz = -INFINITY, sm_ = 0.f;
        int am = 0;
        for (int p = lane; p < P; p += 32) {
            const float v = z[p * Tp + t];
            sz += v;
            if (v > mz) { mz = v; am = p; }
            sm_ += __expf(v - rmax[p]) * rinv[p];
        }
        sz = warp_sum(sz);
        sm_ = warp_sum(sm_);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // arg-max: ties resolve to the lowest pixel index (torch.max)
            const float ov = __shfl_xor_sync(0xffffffffu, mz, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (ov > mz || (ov == mz && oa < am)) { mz = ov; am = oa; }
        }
        if (lane == 0) {
            const float m = sm_ / P;
            cls_out[static_cast<long>(b) * T + t] = sz / P + mz + __powf(1.f - m, focal_p) * __logf(focal_l + m);
            mbar_out[static_cast<long>(b) * T + t] = m;
            argmax_out[static_cast<long>(b) * T + t] = am;
            if (t == b) cls_fg[b] = m;
        }
    }
}

// D bf16 [B,P,Tp] = dL/dR ; dlogit_scale[b] = per-image sum dz * z (fp32 [B]; the caller adds them in order: no atomics).
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ R, const float* __restrict__ logit_scale,
                                                       const float* __restrict__ dcls_out, const float* __restrict__ dcls_fg,
                                                       const float* __restrict__ dmaps, const float* __restrict__ mbar,
                                                       const int* __restrict__ argmax, __nv_bfloat16* __restrict__ D,
                                                       float* __restrict__ dlogit_scale, int P, int T, int Tp, float focal_p,
                                                       float focal_l) {
    extern __shared__ float sm[];   // z [P*Tp], rmax [P], rinv [P], inner [P], G [T], dc [T], am [T]
    float* z = sm;
    float* rmax = sm + P * Tp;
    float* rinv = rmax + P;
    float* inner = rinv + P;
    float* G = inner + P;
    float* dc = G + T;
    int* am = reinterpret_cast<int*>(dc + T);
    __shared__ float red[8];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float es = __expf(*logit_scale);
    const long base = static_cast<long>(b) * P * Tp;
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) z[i] = R[base + i] * es;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const float m = mbar[static_cast<long>(b) * T + t];
        const float g = dcls_out != nullptr ? dcls_out[static_cast<long>(b) * T + t] : 0.f;
        // f(m) = (1-m)^k log(l+m);  f' = -k (1-m)^(k-1) log(l+m) + (1-m)^k / (l+m)
        const float om = 1.f - m;
        const float fprime = -focal_p * __powf(om, focal_p - 1.f) * __logf(focal_l + m) + __powf(om, focal_p) / (focal_l + m);
        float gg = g * fprime;
        if (t == b && dcls_fg != nullptr) gg += dcls_fg[b];
        G[t] = gg / P;
        dc[t] = g;
        am[t] = argmax[static_cast<long>(b) * T + t];
    }
    __syncthreads();
    for (int p = warp; p < P; p += nw) {
        float mx = 1.f;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, z[p * Tp + t]);
        mx = warp_max(mx);
        float s = 0.f, in = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float e = __expf(z[p * Tp + t] - mx);
            s += e;
            in += e * G[t];
        }
        s = warp_sum(s) + __expf(1.f - mx);
        in = warp_sum(in);
        if (lane == 0) { rmax[p] = mx; rinv[p] = 1.f / s; inner[p] = in / s; }
    }
    __syncthreads();
    float acc = 0.f;
    for (int i = threadIdx.x; i < P * Tp; i += blockDim.x) {
        const int p = i / Tp, t = i - p * Tp;
        float dz = 0.f;
        if (t < T) {
            const float m = __expf(z[i] - rmax[p]) * rinv[p];
            dz = dc[t] * (1.f / P + (am[t] == p ? 1.f : 0.f)) + m * (G[t] - inner[p]);
            if (t == b && dmaps != nullptr) dz += dmaps[static_cast<long>(b) * P + p];
            acc += dz * z[i];
        }
        D[base + i] = __float2bfloat16(dz * es);
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && dlogit_scale != nullptr) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += red[w];
        dlogit_scale[b] = s;          // per-image partial (plain store); the caller adds the B values in order
    }
}

// ------------------------------------------------------------------------------------------------ bilinear up-sampling
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, bool align, int& i0, int& i1, float& w) {
    float s = align ? scale * dst : fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = min(static_cast<int>(s), in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    w = s - i0;
}

// maps f32 [B, h, w] -> relu / sig f32 [B, 1, H, W]  (align_corners = False).  thread = 4 consecutive x.
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ maps, float* __restrict__ relu_out,
                                                           float* __restrict__ sig_out, int B, int h, int w, int H, int W) {
    const long total = static_cast<long>(B) * H * (W / 4);
    const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int x4 = static_cast<int>(i % (W / 4));
        const long r = i / (W / 4);
        const int y = static_cast<int>(r % H), b = static_cast<int>(r / H);
        int y0, y1;
        float wy;
        src_index(y, sy, h, false, y0, y1, wy);
        const float* m = maps + static_cast<long>(b) * h * w;
        float v[4], sg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x0, x1;
            float wx;
            src_index(x4 * 4 + k, sx, w, false, x0, x1, wx);
            const float top = m[y0 * w + x0] * (1.f - wx) + m[y0 * w + x1] * wx;
            const float bot = m[y1 * w + x0] * (1.f - wx) + m[y1 * w + x1] * wx;
            const float s = top * (1.f - wy) + bot * wy;
            v[k] = fmaxf(s, 0.f);
            sg[k] = 1.f / (1.f + __expf(-s));
        }
        const long o = (static_cast<long>(b) * H + y) * W + x4 * 4;
        if (relu_out != nullptr) *reinterpret_cast<float4*>(relu_out + o) = make_float4(v[0], v[1], v[2], v[3]);
        if (sig_out != nullptr) *reinterpret_cast<float4*>(sig_out + o) = make_float4(sg[0], sg[1], sg[2], sg[3]);
    }
}

// dmaps[b, ry, rx] = sum over the output pixels that read logit (ry,rx) of (dsig * sig(1-sig) + drelu * [seg>0]) * wy * wx.
// Separable gather, no atomics: one CTA per (b, logit row ry).  Phase 1: thread x folds the <= 3 * H/h output rows that read
// logit row ry into col[x] (coalesced row reads, every output row is read by the two CTAs it interpolates between); phase 2:
// one warp per logit column folds col[] with the x weights.  (The per-logit window form visited every output pixel ~9 times
// with two index computations each: 135 us at batch 48, the largest kernel of the head.)  sig = saved forward output
// (seg > 0 <=> sig > 0.5).
__global__ void __launch_bounds__(320) upsample_bwd_kernel(const float* __restrict__ drelu, const float* __restrict__ dsig,
                                                           const float* __restrict__ sig, float* __restrict__ dmaps, int h, int w, int H,
                                                           int W) {
    extern __shared__ float col[];   // [W]
    const int b = blockIdx.y, ry = blockIdx.x;
    const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
    const int ry_span = (H + h - 1) / h, rx_span = (W + w - 1) / w;
    const int ylo = max(0, (ry - 1) * ry_span - 1), yhi = min(H, (ry + 2) * ry_span + 1);
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        float acc = 0.f;
#pragma unroll 4
        for (int y = ylo; y < yhi; ++y) {
            int y0, y1;
            float wy;
            src_index(y, sy, h, false, y0, y1, wy);
            const float cy = (y0 == ry ? 1.f - wy : 0.f) + (y1 == ry ? wy : 0.f);
            if (cy != 0.f) {
                const long o = (static_cast<long>(b) * H + y) * W + x;
                const float sg = sig[o];
                float g = 0.f;
                if (dsig != nullptr) g += dsig[o] * sg * (1.f - sg);
                if (drelu != nullptr && sg > 0.5f) g += drelu[o];
                acc = fmaf(g, cy, acc);
            }
        }
        col[x] = acc;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int rx = warp; rx < w; rx += nwarps) {
        const int xlo = max(0, (rx - 1) * rx_span - 1), xhi = min(W, (rx + 2) * rx_span + 1);
        float a = 0.f;
        for (int x = xlo + lane; x < xhi; x += 32) {
            int x0, x1;
            float wx;
            src_index(x, sx, w, false, x0, x1, wx);
            const float cx = (x0 == rx ? 1.f - wx : 0.f) + (x1 == rx ? wx : 0.f);
            a = fmaf(col[x], cx, a);
        }
        a = warp_sum(a);
        if (lane == 0) dmaps[(static_cast<long>(b) * h + ry) * w + rx] = a;
    }
}

// ------------------------------------------------------------------------------------------------ mask-and-resize (K9)
__device__ __forceinline__ float bilerp(const float* __restrict__ src, int W, int y0, int y1, float wy, int x0, int x1, float wx) {
    const float top = __ldg(src + static_cast<long>(y0) * W + x0) * (1.f - wx) + __ldg(src + static_cast<long>(y0) * W + x1) * wx;
    const float bot = __ldg(src + static_cast<long>(y1) * W + x0) * (1.f - wx) + __ldg(src + static_cast<long>(y1) * W + x1) * wx;
    return top * (1.f - wy) + bot * wy;
}

// patches bf16 [B * (O/ps)^2, 3*ps*ps], k = c*ps*ps + py*ps + px  <-  cam(y,x) * img_c(y,x), both resampled S -> O with
// align_corners=True (identity when S == O).  sig may be NULL (plain patchify of img).  thread = 8 consecutive px.
__global__ void __launch_bounds__(256) mask_resize_fwd_kernel(const float* __restrict__ sig, const float* __restrict__ img,
                                                              __nv_bfloat16* __restrict__ patches, float* __restrict__ fg_out, int B,
                                                              int S, int O, int ps) {
    const int G = O / ps, X8 = O / 8;
    const long total = static_cast<long>(B) * O * X8;
    const float sc = O > 1 ? static_cast<float>(S - 1) / (O - 1) : 0.f;
    const bool same = S == O;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int x8 = static_cast<int>(i % X8);
        const long r = i / X8;
        const int y = static_cast<int>(r % O), b = static_cast<int>(r / O);
        int y0 = y, y1 = y;
        float wy = 0.f;
        if (!same) src_index(y, sc, S, true, y0, y1, wy);
        float cam[8], v[3][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int x = x8 * 8 + k;
            int x0 = x, x1 = x;
            float wx = 0.f;
            if (!same) src_index(x, sc, S, true, x0, x1, wx);
            cam[k] = sig != nullptr ? bilerp(sig + static_cast<long>(b) * S * S, S, y0, y1, wy, x0, x1, wx) : 1.f;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[c][k] = cam[k] * bilerp(img + (static_cast<long>(b) * 3 + c) * S * S, S, y0, y1, wy, x0, x1, wx);
        }
        const int gy = y / ps, py = y % ps, gx = (x8 * 8) / ps, px = (x8 * 8) % ps;
        const long prow = (static_cast<long>(b) * G + gy) * G + gx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (patches != nullptr) {
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[c][2 * j], v[c][2 * j + 1]);
                *reinterpret_cast<uint4*>(patches + prow * (3 * ps * ps) + (c * ps + py) * ps + px) = o;
            }
            if (fg_out != nullptr) {
                float* f = fg_out + ((static_cast<long>(b) * 3 + c) * O + y) * O + x8 * 8;
                *reinterpret_cast<float4*>(f) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
                *reinterpret_cast<float4*>(f + 4) = make_float4(v[c][4], v[c][5], v[c][6], v[c][7]);
            }
        }
    }
}

// dcam f32 [B, O, O] = sum_c dpatch(b,c,y,x) * img_O(b,c,y,x)
__global__ void __launch_bounds__(256) mask_resize_dcam_kernel(const __nv_bfloat16* __restrict__ dpatches, const float* __restrict__ img,
                                                               float* __restrict__ dcam, int B, int S, int O, int ps) {
    const int G = O / ps, X8 = O / 8;
    const long total = static_cast<long>(B) * O * X8;
    const float sc = O > 1 ? static_cast<float>(S - 1) / (O - 1) : 0.f;
    const bool same = S == O;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int x8 = static_cast<int>(i % X8);
        const long r = i / X8;
        const int y = static_cast<int>(r % O), b = static_cast<int>(r / O);
        int y0 = y, y1 = y;
        float wy = 0.f;
        if (!same) src_index(y, sc, S, true, y0, y1, wy);
        const int gy = y / ps, py = y % ps, gx = (x8 * 8) / ps, px = (x8 * 8) % ps;
        const long prow = (static_cast<long>(b) * G + gy) * G + gx;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint4 dv = *reinterpret_cast<const uint4*>(dpatches + prow * (3 * ps * ps) + (c * ps + py) * ps + px);
            const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = x8 * 8 + k;
                int x0 = x, x1 = x;
                float wx = 0.f;
                if (!same) src_index(x, sc, S, true, x0, x1, wx);
                const float2 df = __bfloat1622float2(d2[k >> 1]);
                acc[k] += ((k & 1) ? df.y : df.x) * bilerp(img + (static_cast<long>(b) * 3 + c) * S * S, S, y0, y1, wy, x0, x1, wx);
            }
        }
        float* o = dcam + (static_cast<long>(b) * O + y) * O + x8 * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// dsig f32 [B, S, S]: adjoint of the align_corners=True S -> O resampling, gather form (each source pixel is read by
// at most ceil(O/S)+1 outputs per axis).
__global__ void __launch_bounds__(256) mask_resize_dsig_kernel(const float* __restrict__ dcam, float* __restrict__ dsig, int B, int S,
                                                               int O) {
    const long total = static_cast<long>(B) * S * S;
    const float sc = O > 1 ? static_cast<float>(S - 1) / (O - 1) : 0.f;
    const float inv = S > 1 ? static_cast<float>(O - 1) / (S - 1) : 0.f;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(i % S);
        const long r = i / S;
        const int Y = static_cast<int>(r % S), b = static_cast<int>(r / S);
        if (S == O) { dsig[i] = dcam[i]; continue; }
        const int ylo = max(0, static_cast<int>(floorf((Y - 1) * inv)) - 1), yhi = min(O - 1, static_cast<int>(ceilf((Y + 1) * inv)) + 1);
        const int xlo = max(0, static_cast<int>(floorf((X - 1) * inv)) - 1), xhi = min(O - 1, static_cast<int>(ceilf((X + 1) * inv)) + 1);
        float acc = 0.f;
        for (int y = ylo; y <= yhi; ++y) {
            int y0, y1;
            float wy;
            src_index(y, sc, S, true, y0, y1, wy);
            const float cy = (y0 == Y ? 1.f - wy : 0.f) + (y1 == Y ? wy : 0.f);
            if (cy == 0.f) continue;
            for (int x = xlo; x <= xhi; ++x) {
                int x0, x1;
                float wx;
                src_index(x, sc, S, true, x0, x1, wx);
                const float cx = (x0 == X ? 1.f - wx : 0.f) + (x1 == X ? wx : 0.f);
                if (cx != 0.f) acc += cy * cx * dcam[(static_cast<long>(b) * O + y) * O + x];
            }
        }
        dsig[i] = acc;
    }
}

// plain bilinear resize, align_corners=True: src f32 [NC, H, W] -> dst f32 [NC, OH, OW]  (validate.py:180, demo.py:94)
__global__ void __launch_bounds__(256) resize_ac_kernel(const float* __restrict__ src, float* __restrict__ dst, long NC, int H, int W,
                                                        int OH, int OW) {
    const long total = NC * OH * OW;
    const float sy = OH > 1 ? static_cast<float>(H - 1) / (OH - 1) : 0.f, sx = OW > 1 ? static_cast<float>(W - 1) / (OW - 1) : 0.f;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % OW);
        const long r = i / OW;
        const int y = static_cast<int>(r % OH);
        const long nc = r / OH;
        int y0, y1, x0, x1;
        float wy, wx;
        src_index(y, sy, H, true, y0, y1, wy);
        src_index(x, sx, W, true, x0, x1, wx);
        dst[i] = bilerp(src + nc * H * W, W, y0, y1, wy, x0, x1, wx);
    }
}

// ------------------------------------------------------------------------------------------------ losses (K12)
// f bf16 [B, D] image features, g bf16 [B*(1+K), D] text features (positives first, then negatives b-major),
// cls f32 [B, B].   out[4] = {loss, l1, l4, l5}.   Single CTA, warp per sample: deterministic.
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }

template <typename T>
struct LossParams {
    const T* f;
    const T* g;
    const float* cls;
    float* out;
    const float* dout;          // backward: gradient w.r.t. out[4]
    T* df;                      // backward: [B, D]
    float* dcls;                // backward: [B, B]
    int B, D, K;
    float w1, w4, w5;
};

__device__ __forceinline__ float softplus(float x) { return x > 15.f ? x : log1pf(__expf(x)); }

template <bool kBackward, typename T>
__global__ void __launch_bounds__(512) loss_kernel(const LossParams<T> p) {
    __shared__ float part[3][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int B = p.B, D = p.D, K = p.K;
    float c1 = 0.f, c4 = 0.f, c5 = 0.f;
    if (kBackward) {
        c1 = p.dout[0] * p.w1 + p.dout[1];
        c4 = p.dout[0] * p.w4 + p.dout[2];
        c5 = p.dout[0] * p.w5 + p.dout[3];
    }
    float l1 = 0.f, l4 = 0.f, l5 = 0.f;
    for (int b = warp; b < B; b += nw) {
        const T* fr = p.f + static_cast<long>(b) * D;
        float ff = 0.f;
        for (int i = lane; i < D; i += 32) { const float v = ldf(fr + i); ff += v * v; }
        const float finv = rsqrtf(warp_sum(ff));
        // positive text
        const T* gr = p.g + static_cast<long>(b) * D;
        float gg = 0.f, fg = 0.f;
        for (int i = lane; i < D; i += 32) {
            const float gv = ldf(gr + i);
            gg += gv * gv;
            fg += gv * ldf(fr + i);
        }
        const float ginv = rsqrtf(warp_sum(gg));
        const float cosv = warp_sum(fg) * finv * ginv;
        const float cl = fminf(fmaxf(cosv, 1e-4f), 0.9999f);
        l1 += -__logf(cl);
        // coefficient of the unit text vectors in dL/dfn
        float coef_pos = 0.f;
        if (kBackward && cosv > 1e-4f && cosv < 0.9999f) coef_pos = -c1 / (B * cosv);
        float coef_neg[8], ninv[8];
        for (int k = 0; k < K; ++k) {
            const T* nr = p.g + (static_cast<long>(B) + static_cast<long>(b) * K + k) * D;
            float nn = 0.f, fn = 0.f;
            for (int i = lane; i < D; i += 32) {
                const float nv = ldf(nr + i);
                nn += nv * nv;
                fn += nv * ldf(fr + i);
            }
            ninv[k] = rsqrtf(warp_sum(nn));
            const float cn = warp_sum(fn) * finv * ninv[k];
            l5 += -__logf(1.f - cn) / K;
            coef_neg[k] = kBackward ? c5 / (static_cast<float>(B) * K * (1.f - cn)) : 0.f;
        }
        if (kBackward) {
            // dfn = coef_pos * gn + sum_k coef_neg[k] * nn_k ;  df = (dfn - fn <fn, dfn>) * finv
            float dot = 0.f;
            for (int i = lane; i < D; i += 32) {
                float d = coef_pos * ginv * ldf(gr + i);
                for (int k = 0; k < K; ++k)
                    d += coef_neg[k] * ninv[k] * ldf(p.g + (static_cast<long>(B) + static_cast<long>(b) * K + k) * D + i);
                dot += d * ldf(fr + i) * finv;
            }
            dot = warp_sum(dot);
            for (int i = lane; i < D; i += 32) {
                float d = coef_pos * ginv * ldf(gr + i);
                for (int k = 0; k < K; ++k)
                    d += coef_neg[k] * ninv[k] * ldf(p.g + (static_cast<long>(B) + static_cast<long>(b) * K + k) * D + i);
                stf(p.df + static_cast<long>(b) * D + i, (d - ldf(fr + i) * finv * dot) * finv);
            }
        }
        // multilabel soft margin row b (labels = identity): -[y log s(c) + (1-y) log s(-c)] = softplus(-c) | softplus(c)
        float r4 = 0.f;
        for (int j = lane; j < B; j += 32) {
            const float c = p.cls[static_cast<long>(b) * B + j];
            r4 += j == b ? softplus(-c) : softplus(c);
            if (kBackward) {
                const float s = 1.f / (1.f + __expf(-c));
                p.dcls[static_cast<long>(b) * B + j] = c4 * (j == b ? s - 1.f : s) / (static_cast<float>(B) * B);
            }
        }
        l4 += warp_sum(r4);
    }
    if (kBackward) return;
    if (lane == 0) { part[0][warp] = l1; part[1][warp] = l4; part[2][warp] = l5; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s1 = 0.f, s4 = 0.f, s5 = 0.f;
        for (int w = 0; w < nw; ++w) { s1 += part[0][w]; s4 += part[1][w]; s5 += part[2][w]; }
        s1 /= B;
        s4 /= static_cast<float>(B) * B;
        s5 /= B;
        p.out[0] = p.w1 * s1 + p.w4 * s4 + p.w5 * s5;
        p.out[1] = s1;
        p.out[2] = s4;
        p.out[3] = s5;
    }
}

int grid_for(long n, int threads) {
    long g = (n + threads - 1) / threads;
    const long cap = static_cast<long>(tris::sm_count()) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}


// ------------------------------------------------------------------------------------------------ PRMS (validate.py:304-332)
// get_scores (validate.py:120-127) for the S candidate maps of a ref against its S sentences, summed over the sentences, and
// the arg-max (first maximum wins, like the reference's `score > max_info['score']`).  f [S, D] image features of fg_j,
// g [S, D] text features; one CTA, warp w handles candidates w, w + nwarps, ...
__global__ void __launch_bounds__(256) prms_select_kernel(const __nv_bfloat16* __restrict__ f, const __nv_bfloat16* __restrict__ g,
                                                          float* __restrict__ scores, int* __restrict__ best, int S, int D) {
    __shared__ float s_ng[32], s_sc[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < S; j += nw) {
        float q = 0.f;
        for (int d = lane; d < D; d += 32) { const float x = __bfloat162float(g[j * D + d]); q = fmaf(x, x, q); }
        q = warp_sum(q);
        if (lane == 0) s_ng[j] = rsqrtf(q);
    }
    __syncthreads();
    for (int i = warp; i < S; i += nw) {
        float q = 0.f;
        for (int d = lane; d < D; d += 32) { const float x = __bfloat162float(f[i * D + d]); q = fmaf(x, x, q); }
        const float nf = rsqrtf(warp_sum(q));
        float tot = 0.f;
        for (int j = 0; j < S; ++j) {
            float dot = 0.f;
            for (int d = lane; d < D; d += 32) dot = fmaf(__bfloat162float(f[i * D + d]), __bfloat162float(g[j * D + d]), dot);
            tot += warp_sum(dot) * nf * s_ng[j];
        }
        if (lane == 0) s_sc[i] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0;
        for (int i = 0; i < S; ++i) {
            if (scores != nullptr) scores[i] = s_sc[i];
            if (s_sc[i] > s_sc[b]) b = i;
        }
        *best = b;
    }
}

// validate.py:183-191 for one map: cam /= max + 1e-5 ; pred = cam > 1e-9 ; I = |pred & target|, U = |pred | target| ; pointing
// game: does the (first) arg-max fall on the target?  cams [S, H*W] with *sel choosing the map (PRMS) or sel == nullptr -> map 0.
// out [H*W] normalised map ; stats[4] = I, U, hit, max.  One CTA of 1024 threads (the map is L2-resident).
__global__ void __launch_bounds__(1024) cam_metrics_kernel(const float* __restrict__ cams, const int* __restrict__ sel,
                                                           const long long* __restrict__ target, float* __restrict__ out,
                                                           float* __restrict__ stats, int HW) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ float s_a[32], s_b[32];
    const float* cam = cams + static_cast<long>(sel != nullptr ? *sel : 0) * HW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mv = -3.4e38f;
    int mi = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float v = cam[i];
        if (v > mv) { mv = v; mi = i; }         // strided scan: indices increase, so the first maximum of this thread is kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
    }
    if (lane == 0) { s_v[warp] = mv; s_i[warp] = mi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (s_v[w] > s_v[0] || (s_v[w] == s_v[0] && s_i[w] < s_i[0])) { s_v[0] = s_v[w]; s_i[0] = s_i[w]; }
    }
    __syncthreads();
    const float inv = 1.f / (s_v[0] + 1e-5f);
    float ii = 0.f, uu = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float v = cam[i] * inv;
        out[i] = v;
        const bool p = v > 1e-9f, t = target[i] != 0;
        ii += (p && t) ? 1.f : 0.f;
        uu += (p || t) ? 1.f : 0.f;
    }
    ii = warp_sum(ii); uu = warp_sum(uu);
    if (lane == 0) { s_a[warp] = ii; s_b[warp] = uu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 32; ++w) { a += s_a[w]; b += s_b[w]; }
        stats[0] = a; stats[1] = b; stats[2] = target[s_i[0]] != 0 ? 1.f : 0.f; stats[3] = s_v[0];
    }
}

}  // namespace

extern "C" {

int tris_xattn_softmax_fwd(const float* S1, const float* S2T, void* PA, void* PAc, void* PTt, int B, int P, int T, int Tp, float scale,
                           tris_stream_t stream) {
    const size_t smem = (static_cast<size_t>(P) * Tp * 2 + Tp) * 4;
    if (T < 1 || Tp < T || Tp % 8) return tris::fail(TRIS_ERR_SHAPE, "xattn_softmax: T=%d Tp=%d", T, Tp);
    if (smem > 200 * 1024) return tris::fail(TRIS_ERR_SHAPE, "xattn_softmax: P*Tp=%d*%d exceeds the shared-memory tile", P, Tp);
    if (smem > 48 * 1024) TRIS_CUDA_OK(cudaFuncSetAttribute(xattn_softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_softmax_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(S1, S2T, (__nv_bfloat16*)PA, (__nv_bfloat16*)PAc, (__nv_bfloat16*)PTt, P, T, Tp,
                                                                         scale);
    TRIS_LAUNCH_OK("xattn_softmax_fwd");
    return TRIS_OK;
}

int tris_xattn_softmax_bwd(const void* PA, const float* dPA, const void* PTt, const float* dPTt, void* dS1, void* dS2T, int B, int P,
                           int T, int Tp, float scale, tris_stream_t stream) {
    if (T < 1 || Tp < T || Tp % 8) return tris::fail(TRIS_ERR_SHAPE, "xattn_softmax_bwd: T=%d Tp=%d", T, Tp);
    xattn_softmax_bwd_kernel<<<B, 256, 2 * T * 4, (cudaStream_t)stream>>>((const __nv_bfloat16*)PA, dPA, (const __nv_bfloat16*)PTt, dPTt,
                                                                    (__nv_bfloat16*)dS1, (__nv_bfloat16*)dS2T, P, T, Tp, scale);
    TRIS_LAUNCH_OK("xattn_softmax_bwd");
    return TRIS_OK;
}

int tris_bcast_mix(const void* base, const void* x, void* out, long per, int B, float a, tris_stream_t stream) {
    if (per % 8) return tris::fail(TRIS_ERR_ALIGN, "bcast_mix: per=%ld %% 8", per);
    const long total8 = per / 8 * B;
    bcast_mix_kernel<<<grid_for(total8, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)base, (const __nv_bfloat16*)x,
                                                                             (__nv_bfloat16*)out, per / 8, total8, a);
    TRIS_LAUNCH_OK("bcast_mix");
    return TRIS_OK;
}

int tris_batch_sum(const void* x, const void* add, void* out, long per, int B, float a, tris_stream_t stream) {
    batch_sum_kernel<<<static_cast<int>((per + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)add,
                                                                                           (__nv_bfloat16*)out, per, B, a);
    TRIS_LAUNCH_OK("batch_sum");
    return TRIS_OK;
}

int tris_relu_mask(const float* g, const void* y, void* dst, long n, tris_stream_t stream) {
    relu_mask_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)y, (__nv_bfloat16*)dst, n);
    TRIS_LAUNCH_OK("relu_mask");
    return TRIS_OK;
}

int tris_head_fwd(const float* R, const float* logit_scale, float* cls_out, float* cls_fg, float* maps, float* mbar, int* argmax,
                  float* es_out, int B, int P, int T, int Tp, float focal_p, float focal_l, int train, tris_stream_t stream) {
    if (T != B) return tris::fail(TRIS_ERR_SHAPE, "head_fwd: the diagonal map needs T == B (got T=%d B=%d)", T, B);
    const size_t smem = (static_cast<size_t>(P) * Tp + 2 * P) * 4;
    if (smem > 200 * 1024) return tris::fail(TRIS_ERR_SHAPE, "head_fwd: P*Tp too large");
    if (smem > 48 * 1024) TRIS_CUDA_OK(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(R, logit_scale, cls_out, cls_fg, maps, mbar, argmax, es_out, P, T, Tp, focal_p,
                                                           focal_l, train);
    TRIS_LAUNCH_OK("head_fwd");
    return TRIS_OK;
}

int tris_head_bwd(const float* R, const float* logit_scale, const float* dcls_out, const float* dcls_fg, const float* dmaps,
                  const float* mbar, const int* argmax, void* D, float* dlogit_scale, int B, int P, int T, int Tp, float focal_p,
                  float focal_l, tris_stream_t stream) {
    const size_t smem = (static_cast<size_t>(P) * Tp + 3 * P + 3 * T) * 4;
    if (smem > 200 * 1024) return tris::fail(TRIS_ERR_SHAPE, "head_bwd: P*Tp too large");
    if (smem > 48 * 1024) TRIS_CUDA_OK(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_bwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(R, logit_scale, dcls_out, dcls_fg, dmaps, mbar, argmax, (__nv_bfloat16*)D,
                                                           dlogit_scale, P, T, Tp, focal_p, focal_l);
    TRIS_LAUNCH_OK("head_bwd");
    return TRIS_OK;
}

int tris_upsample_fwd(const float* maps, float* relu_out, float* sig_out, int B, int h, int w, int H, int W, tris_stream_t stream) {
    if (W % 4) return tris::fail(TRIS_ERR_ALIGN, "upsample_fwd: W=%d %% 4", W);
    upsample_fwd_kernel<<<grid_for(static_cast<long>(B) * H * (W / 4), 256), 256, 0, (cudaStream_t)stream>>>(maps, relu_out, sig_out, B, h, w,
                                                                                                            H, W);
    TRIS_LAUNCH_OK("upsample_fwd");
    return TRIS_OK;
}

int tris_upsample_bwd(const float* drelu, const float* dsig, const float* sig, float* dmaps, int B, int h, int w, int H, int W,
                      tris_stream_t stream) {
    upsample_bwd_kernel<<<dim3(h, B), 320, W * sizeof(float), (cudaStream_t)stream>>>(drelu, dsig, sig, dmaps, h, w, H, W);
    TRIS_LAUNCH_OK("upsample_bwd");
    return TRIS_OK;
}

int tris_mask_resize_fwd(const float* sig, const float* img, void* patches, float* fg, int B, int S, int O, int ps, tris_stream_t stream) {
    if (O % ps || ps % 8) return tris::fail(TRIS_ERR_SHAPE, "mask_resize: O=%d ps=%d", O, ps);
    mask_resize_fwd_kernel<<<grid_for(static_cast<long>(B) * O * (O / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        sig, img, (__nv_bfloat16*)patches, fg, B, S, O, ps);
    TRIS_LAUNCH_OK("mask_resize_fwd");
    return TRIS_OK;
}

int tris_mask_resize_bwd(const void* dpatches, const float* img, float* dcam, float* dsig, int B, int S, int O, int ps,
                         tris_stream_t stream) {
    if (O % ps || ps % 8) return tris::fail(TRIS_ERR_SHAPE, "mask_resize_bwd: O=%d ps=%d", O, ps);
    mask_resize_dcam_kernel<<<grid_for(static_cast<long>(B) * O * (O / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dpatches, img, dcam, B, S, O, ps);
    TRIS_LAUNCH_OK("mask_resize_dcam");
    mask_resize_dsig_kernel<<<grid_for(static_cast<long>(B) * S * S, 256), 256, 0, (cudaStream_t)stream>>>(dcam, dsig, B, S, O);
    TRIS_LAUNCH_OK("mask_resize_dsig");
    return TRIS_OK;
}

int tris_prms_select(const void* f, const void* g, float* scores, int* best, int S, int D, tris_stream_t stream) {
    if (S < 1 || S > 32) return tris::fail(TRIS_ERR_SHAPE, "tris_prms_select: 1 <= S <= 32 sentences (got %d)", S);
    prms_select_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)f, (const __nv_bfloat16*)g, scores, best, S, D);
    TRIS_LAUNCH_OK("prms_select_kernel");
    return TRIS_OK;
}

int tris_cam_metrics(const float* cams, const int* sel, const long long* target, float* out, float* stats, int HW, tris_stream_t stream) {
    cam_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cams, sel, target, out, stats, HW);
    TRIS_LAUNCH_OK("cam_metrics_kernel");
    return TRIS_OK;
}

int tris_resize_bilinear_ac(const float* src, float* dst, long nc, int H, int W, int OH, int OW, tris_stream_t stream) {
    resize_ac_kernel<<<grid_for(nc * OH * OW, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, nc, H, W, OH, OW);
    TRIS_LAUNCH_OK("resize_bilinear_ac");
    return TRIS_OK;
}

int tris_stage1_loss_fwd(const void* f, const void* g, const float* cls, float* out, int B, int D, int K, float w1, float w4, float w5,
                         tris_stream_t stream) {
    if (K > 8) return tris::fail(TRIS_ERR_SHAPE, "stage1_loss: at most 8 negatives per sample (got %d)", K);
    LossParams<__nv_bfloat16> p{(const __nv_bfloat16*)f, (const __nv_bfloat16*)g, cls, out, nullptr, nullptr, nullptr, B, D, K, w1, w4, w5};
    loss_kernel<false, __nv_bfloat16><<<1, 512, 0, (cudaStream_t)stream>>>(p);
    TRIS_LAUNCH_OK("stage1_loss_fwd");
    return TRIS_OK;
}

/* fp32-feature variant of the forward (fp32 parity mode) */
int tris_stage1_loss_fwd_f32(const float* f, const float* g, const float* cls, float* out, int B, int D, int K, float w1, float w4, float w5,
                             tris_stream_t stream) {
    if (K > 8) return tris::fail(TRIS_ERR_SHAPE, "stage1_loss: at most 8 negatives per sample (got %d)", K);
    LossParams<float> p{f, g, cls, out, nullptr, nullptr, nullptr, B, D, K, w1, w4, w5};
    loss_kernel<false, float><<<1, 512, 0, (cudaStream_t)stream>>>(p);
    TRIS_LAUNCH_OK("stage1_loss_fwd_f32");
    return TRIS_OK;
}

int tris_stage1_loss_bwd(const void* f, const void* g, const float* cls, const float* dout, void* df, float* dcls, int B, int D, int K,
                         float w1, float w4, float w5, tris_stream_t stream) {
    if (K > 8) return tris::fail(TRIS_ERR_SHAPE, "stage1_loss: at most 8 negatives per sample (got %d)", K);
    LossParams<__nv_bfloat16> p{(const __nv_bfloat16*)f, (const __nv_bfloat16*)g, cls, nullptr, dout, (__nv_bfloat16*)df, dcls, B, D, K, w1, w4, w5};
    loss_kernel<true, __nv_bfloat16><<<1, 512, 0, (cudaStream_t)stream>>>(p);
    TRIS_LAUNCH_OK("stage1_loss_bwd");
    return TRIS_OK;
}

}  // extern "C"
