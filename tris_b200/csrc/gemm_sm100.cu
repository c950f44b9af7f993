// tcgen05 / TMA / TMEM GEMM + implicit-GEMM convolution for sm_100a (B200).
//
// One persistent, warp-specialised kernel (one CTA per SM):
//   warp 0  : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1  : UMMA issuer    (tcgen05.mma kind::f16, 128 x BN x 16, fp32 accumulators in TMEM, double buffered)
//   warps 2-5: epilogue      (tcgen05.ld -> bias / activation / residual / BN column statistics -> global)
// The same kernel serves linear layers, 1x1 convs (plain GEMM on NHWC), 3x3 convs (forward and data-gradient:
// A read through a 4-D TMA box with zero-filled halo, one k-block per (tap, 64 channels)) and weight gradients
// (both operands MN-major, contraction over rows / pixels, split-K with fp32 reductions).
//
// Replaces cuDNN/cuBLAS calls behind CLIP/clip/model.py:17-40 (Bottleneck convs), :366-378 (transformer linears),
// model/model_stage1.py:36-37 and model/attn.py:69-109 of the reference.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator buffers

struct KParams {
    int M, N, K;
    int a_mode, b_mode, wgrad, flip, taps;
    int tiles_m, tiles_n, tiles_tap, split_k, kblocks, kb_per_split;
    int bk, n_mma, bn;
    int img_n, img_h, img_w, th, tw, tiles_h, tiles_w, cblocks, b_tap_stride;
    uint32_t a_bytes, b_bytes, a_atom, b_atom, tx_bytes, stages;
    uint32_t idesc;
    void* d;
    const float* bias;
    const __nv_bfloat16* residual;
    float* stats;
    __nv_bfloat16* d_pre;            // optional: pre-activation (post-bias) output, bf16 [M, ldd]
    const __nv_bfloat16* dact_src;   // optional: multiply by act'(dact_src[row, col]) instead of applying act
    int ldd, act, out_f32, atomic;
};

struct SmemCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

struct TileCoord {
    int m_t, n_t, tap, split;
    int n_i, h0, w0;  // conv fwd/dgrad output patch
};

__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
    TileCoord c;
    c.n_t = t % p.tiles_n;
    t /= p.tiles_n;
    c.m_t = t % p.tiles_m;
    t /= p.tiles_m;
    c.tap = t % p.tiles_tap;
    c.split = t / p.tiles_tap;
    c.n_i = c.h0 = c.w0 = 0;
    if (p.a_mode == TRIS_OP_CONV && !p.wgrad) {
        int tw_i = c.m_t % p.tiles_w;
        int r = c.m_t / p.tiles_w;
        c.w0 = tw_i * p.tw;
        c.h0 = (r % p.tiles_h) * p.th;
        c.n_i = r / p.tiles_h;
    }
    return c;
}

__device__ __forceinline__ float act_grad(float p, int act) {
    if (act == TRIS_ACT_RELU) return p > 0.f ? 1.f : 0.f;
    if (act == TRIS_ACT_QUICKGELU) {
        const float sg = 1.f / (1.f + __expf(-1.702f * p));
        return sg * (1.f + 1.702f * p * (1.f - sg));
    }
    return 1.f;
}

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == TRIS_ACT_RELU) return fmaxf(v, 0.f);
    if (act == TRIS_ACT_QUICKGELU) return v / (1.f + __expf(-1.702f * v));
    return v;
}

// Sum each of 32 register columns over the 32 lanes of the warp; lane c ends with column c's total in v[0].
__device__ __forceinline__ float warp_column_sums(float (&v)[32], uint32_t lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float send = upper ? v[i] : v[i + off];
            float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(kThreads, 1)
tris_umma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const KParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + p.stages * stage_bytes);
    const uint32_t smem_base = ptx::smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_tap * p.tiles_m * p.tiles_n * p.split_k;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_b);
        for (uint32_t s = 0; s < p.stages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&ctl->full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&ctl->empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(ptx::smem_u32(&ctl->acc_full[b]), 1);
            ptx::mbar_init(ptx::smem_u32(&ctl->acc_empty[b]), 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&ctl->tmem_base), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t);
                const int m0 = c.m_t * kBlockM, n0 = c.n_t * p.bn;
                const int kb0 = c.split * p.kb_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&ctl->empty[stage]), phase ^ 1);
                    const uint32_t bar = ptx::smem_u32(&ctl->full[stage]);
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = sa + p.a_bytes;
                    ptx::mbar_expect_tx(bar, p.tx_bytes);
                    // conv k-block decode (fwd/dgrad: (tap, channel block); wgrad: pixel patch)
                    int tap = 0, cb = kb, dr = 0, ds = 0, pn = 0, ph0 = 0, pw0 = 0;
                    if (p.a_mode == TRIS_OP_CONV) {
                        if (!p.wgrad) {
                            tap = kb / p.cblocks;
                            cb = kb - tap * p.cblocks;
                            if (p.taps == 9) {
                                dr = tap / 3 - 1;
                                ds = tap % 3 - 1;
                                if (p.flip) { dr = -dr; ds = -ds; }
                            }
                        } else {
                            int tw_i = kb % p.tiles_w;
                            int r = kb / p.tiles_w;
                            pw0 = tw_i * p.tw;
                            ph0 = (r % p.tiles_h) * p.th;
                            pn = r / p.tiles_h;
                            if (p.taps == 9) { dr = c.tap / 3 - 1; ds = c.tap % 3 - 1; }
                        }
                    }
                    // ---- A
                    if (p.a_mode == TRIS_OP_K2D) {
                        ptx::tma_load_2d(sa, &map_a, bar, kb * 64, m0);
                    } else if (p.a_mode == TRIS_OP_MN2D) {
                        ptx::tma_load_2d(sa, &map_a, bar, m0, kb * p.bk);
                        ptx::tma_load_2d(sa + p.a_atom, &map_a, bar, m0 + 64, kb * p.bk);
                    } else if (!p.wgrad) {
                        ptx::tma_load_4d(sa, &map_a, bar, cb * 64, c.w0 + ds, c.h0 + dr, c.n_i);
                    } else {
                        ptx::tma_load_4d(sa, &map_a, bar, m0, pw0, ph0, pn);
                        ptx::tma_load_4d(sa + p.a_atom, &map_a, bar, m0 + 64, pw0, ph0, pn);
                    }
                    // ---- B
                    if (p.b_mode == TRIS_OP_K2D) {
                        ptx::tma_load_2d(sb, &map_b, bar, kb * 64, n0);
                    } else if (p.b_mode == TRIS_OP_MN2D) {
                        int inner = n0, outer = kb * p.bk;
                        if (p.a_mode == TRIS_OP_CONV) { inner = tap * p.b_tap_stride + n0; outer = cb * 64; }
                        for (int j = 0; j < p.bn / 64; ++j)
                            ptx::tma_load_2d(sb + j * p.b_atom, &map_b, bar, inner + 64 * j, outer);
                    } else {
                        for (int j = 0; j < p.bn / 64; ++j)
                            ptx::tma_load_4d(sb + j * p.b_atom, &map_b, bar, n0 + 64 * j, pw0 + ds, ph0 + dr, pn);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer (single thread)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            uint32_t acc_phase[2] = {0, 0};
            int it = 0;
            const bool a_mn = (p.a_mode == TRIS_OP_MN2D) || (p.a_mode == TRIS_OP_CONV && p.wgrad);
            const bool b_mn = (p.b_mode != TRIS_OP_K2D);
            const uint32_t a_kstep = a_mn ? 2048u : 32u, b_kstep = b_mn ? 2048u : 32u;
            const uint32_t a_lbo = a_mn ? p.a_atom : 0u, b_lbo = b_mn ? p.b_atom : 0u;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const TileCoord c = decode_tile(p, t);
                const int kb0 = c.split * p.kb_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                const int buf = it & 1;
                ptx::mbar_wait(ptx::smem_u32(&ctl->acc_empty[buf]), acc_phase[buf] ^ 1);
                acc_phase[buf] ^= 1;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * kAccStride;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&ctl->full[stage]), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = sa + p.a_bytes;
                    for (int k = 0; k < p.n_mma; ++k) {
                        const uint64_t da = ptx::umma_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = ptx::umma_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
                        ptx::umma_f16(tmem_d, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(ptx::smem_u32(&ctl->empty[stage]));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(ptx::smem_u32(&ctl->acc_full[buf]));
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        uint32_t acc_phase[2] = {0, 0};
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const TileCoord c = decode_tile(p, t);
            const int buf = it & 1;
            const int n0 = c.n_t * p.bn;
            // output row of this thread
            long grow;
            bool rvalid;
            if (p.a_mode == TRIS_OP_CONV && !p.wgrad) {
                const int hh = c.h0 + row / p.tw, ww = c.w0 + row % p.tw;
                rvalid = (row < p.th * p.tw) && hh < p.img_h && ww < p.img_w;
                grow = (static_cast<long>(c.n_i) * p.img_h + hh) * p.img_w + ww;
            } else {
                grow = static_cast<long>(c.m_t) * kBlockM + row;
                rvalid = grow < p.M;
            }
            const int col_base = (p.wgrad ? c.tap * p.N : 0) + n0;
            ptx::mbar_wait(ptx::smem_u32(&ctl->acc_full[buf]), acc_phase[buf]);
            acc_phase[buf] ^= 1;
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccStride;
            for (int ch = 0; ch < p.bn / 32; ++ch) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(taddr + ch * 32, raw);
                ptx::tmem_ld_wait();
                const int ncol0 = n0 + ch * 32;  // column within N
                if (ncol0 >= p.N) continue;      // warp-uniform
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ncol0 + i < p.N) v[i] += __ldg(p.bias + ncol0 + i);
                }
                if (p.stats != nullptr) {
                    float s1[32], s2[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float x = rvalid ? v[i] : 0.f;
                        s1[i] = x;
                        s2[i] = x * x;
                    }
                    const float c1 = warp_column_sums(s1, lane);
                    const float c2 = warp_column_sums(s2, lane);
                    if (ncol0 + static_cast<int>(lane) < p.N) {
                        atomicAdd(p.stats + ncol0 + lane, c1);
                        atomicAdd(p.stats + p.N + ncol0 + lane, c2);
                    }
                }
                // per-thread predicate from here on: no warp-collective ops inside (tcgen05.ld is .sync.aligned)
                if (rvalid) {
                const long off = grow * p.ldd + col_base + ch * 32;
                if (p.d_pre != nullptr) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (ncol0 + g * 8 < p.N) {
                            uint4 o;
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                o2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                            reinterpret_cast<uint4*>(p.d_pre + off)[g] = o;
                        }
                    }
                }
                if (p.dact_src != nullptr) {
                    const uint4* sp = reinterpret_cast<const uint4*>(p.dact_src + off);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (ncol0 + g * 8 < p.N) {
                            uint4 rr = __ldg(sp + g);
                            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 f = __bfloat1622float2(r2[j]);
                                v[g * 8 + 2 * j] *= act_grad(f.x, p.act);
                                v[g * 8 + 2 * j + 1] *= act_grad(f.y, p.act);
                            }
                        }
                    }
                } else if (p.act != TRIS_ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = act_apply(v[i], p.act);
                }
                if (p.residual != nullptr) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (ncol0 + g * 8 < p.N) {
                            uint4 rr = __ldg(rp + g);
                            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 f = __bfloat1622float2(r2[j]);
                                v[g * 8 + 2 * j] += f.x;
                                v[g * 8 + 2 * j + 1] += f.y;
                            }
                        }
                    }
                }
                if (p.out_f32) {
                    float* dp = reinterpret_cast<float*>(p.d) + off;
                    if (p.atomic) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ncol0 + i < p.N) atomicAdd(dp + i, v[i]);
                    } else {
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (ncol0 + g * 4 < p.N)
                                reinterpret_cast<float4*>(dp)[g] =
                                    make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                    }
                } else {
                    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + off;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (ncol0 + g * 8 < p.N) {
                            uint4 o;
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                o2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                            reinterpret_cast<uint4*>(dp)[g] = o;
                        }
                    }
                }
                }  // rvalid
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->acc_empty[buf]));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace

extern "C" int tris_gemm(const tris_gemm_desc* g, tris_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!g || !g->a || !g->b || !g->d) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: null operand");
    if (g->M <= 0 || g->N <= 0 || g->K <= 0) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: empty extent M=%d N=%d K=%d", g->M, g->N, g->K);
    const int bn = g->block_n;
    if (bn < 32 || bn > 256 || bn % 32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: block_n %d not in {32..256 step 32}", bn);
    if (g->N % 8 || g->ldd % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: N=%d and ldd=%d must be multiples of 8", g->N, g->ldd);
    const bool conv = g->a_mode == TRIS_OP_CONV;
    const bool wgrad = conv && g->wgrad;
    const bool b_mn = g->b_mode != TRIS_OP_K2D;
    if (b_mn && bn % 64) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: MN-major B needs block_n %% 64 == 0");
    if ((g->b_mode == TRIS_OP_CONV) != wgrad) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: B conv mode only for wgrad");
    if (g->atomic && g->out_dtype != TRIS_DT_F32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: atomic needs f32 output");
    if (g->split_k > 1 && !g->atomic) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: split_k needs atomic output");
    if (g->residual && g->out_dtype == TRIS_DT_F32 && g->atomic) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: residual+atomic");

    KParams p{};
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.a_mode = g->a_mode; p.b_mode = g->b_mode; p.wgrad = wgrad; p.flip = g->flip; p.taps = g->taps > 0 ? g->taps : 1;
    p.bn = bn;
    p.img_n = g->img_n; p.img_h = g->img_h; p.img_w = g->img_w; p.th = g->tile_h; p.tw = g->tile_w;
    p.b_tap_stride = g->b_tap_stride;
    p.tiles_tap = 1;
    p.tiles_h = p.tiles_w = 1;
    int a_ch = 0;  // channels of the A tensor in conv mode
    if (conv) {
        if (p.taps != 1 && p.taps != 9) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: taps must be 1 or 9");
        if (p.th <= 0 || p.tw <= 0 || p.th > 256 || p.tw > 256) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: bad conv tile");
        p.tiles_h = ceil_div(p.img_h, p.th);
        p.tiles_w = ceil_div(p.img_w, p.tw);
    }
    if (conv && !wgrad) {
        if (p.th * p.tw > kBlockM) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: conv tile %dx%d > 128 rows", p.th, p.tw);
        if (g->K % (p.taps * 64)) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: conv K=%d not a multiple of taps*64", g->K);
        a_ch = g->K / p.taps;
        p.cblocks = a_ch / 64;
        p.bk = 64;
        p.kblocks = p.taps * p.cblocks;
        p.tiles_m = p.img_n * p.tiles_h * p.tiles_w;
        if ((long)p.img_n * p.img_h * p.img_w != g->M) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: conv M != n*h*w");
    } else if (wgrad) {
        p.bk = p.th * p.tw;
        if (p.bk % 16 || p.bk > 96) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: wgrad patch %dx%d must be %%16 and <= 96 rows", p.th, p.tw);
        a_ch = g->M;
        p.kblocks = p.img_n * p.tiles_h * p.tiles_w;
        p.tiles_m = ceil_div(g->M, kBlockM);
        p.tiles_tap = p.taps;
        if ((long)p.img_n * p.img_h * p.img_w != g->K) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: wgrad K != n*h*w");
    } else {
        p.bk = 64;
        p.kblocks = ceil_div(g->K, 64);
        p.tiles_m = ceil_div(g->M, kBlockM);
    }
    p.n_mma = p.bk / 16;
    p.tiles_n = ceil_div(g->N, bn);
    int split = g->split_k > 1 ? g->split_k : 1;
    if (split > p.kblocks) split = p.kblocks;
    p.kb_per_split = ceil_div(p.kblocks, split);
    p.split_k = ceil_div(p.kblocks, p.kb_per_split);

    const bool a_mn = (g->a_mode == TRIS_OP_MN2D) || wgrad;
    p.a_atom = p.bk * 128;
    p.b_atom = p.bk * 128;
    p.a_bytes = a_mn ? 2 * p.a_atom : kBlockM * 128;
    p.b_bytes = b_mn ? (bn / 64) * p.b_atom : bn * 128;
    // bytes actually delivered per stage (full boxes, OOB elements are zero-filled and counted)
    uint32_t a_tx = p.a_bytes, b_tx = p.b_bytes;
    if (conv && !wgrad) a_tx = p.th * p.tw * 128;
    p.tx_bytes = a_tx + b_tx;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    const uint32_t budget = 227 * 1024 - 1024 - sizeof(SmemCtl) - 64;
    p.stages = budget / stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: stage too large (%u bytes)", stage_bytes);
    const size_t smem_bytes = 1024 + p.stages * stage_bytes + sizeof(SmemCtl) + 64;

    p.idesc = ptx::umma_idesc(1u, a_mn ? 1u : 0u, b_mn ? 1u : 0u, kBlockM, bn);
    p.d = g->d; p.bias = g->bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(g->residual); p.stats = g->stats;
    p.d_pre = reinterpret_cast<__nv_bfloat16*>(g->d_pre);
    p.dact_src = reinterpret_cast<const __nv_bfloat16*>(g->dact_src);
    p.ldd = g->ldd; p.act = g->act; p.out_f32 = g->out_dtype == TRIS_DT_F32; p.atomic = g->atomic;

    // ---- tensor maps
    const CUtensorMap *ma = nullptr, *mb = nullptr;
    if (conv) {
        const int C = a_ch;
        if (C % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: conv channels %d %% 8", C);
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.img_w, (uint64_t)p.img_h, (uint64_t)p.img_n};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.img_w * C * 2, (uint64_t)p.img_h * p.img_w * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
        ma = tris::tensor_map_bf16(g->a, 4, dims, str, box);
    } else if (g->a_mode == TRIS_OP_K2D) {
        if (g->lda % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: lda %d %% 8", g->lda);
        uint64_t dims[2] = {(uint64_t)g->K, (uint64_t)g->M};
        uint64_t str[1] = {(uint64_t)g->lda * 2};
        uint32_t box[2] = {64, kBlockM};
        ma = tris::tensor_map_bf16(g->a, 2, dims, str, box);
    } else {
        if (g->lda % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: lda %d %% 8", g->lda);
        uint64_t dims[2] = {(uint64_t)g->M, (uint64_t)g->K};
        uint64_t str[1] = {(uint64_t)g->lda * 2};
        uint32_t box[2] = {64, (uint32_t)p.bk};
        ma = tris::tensor_map_bf16(g->a, 2, dims, str, box);
    }
    if (!ma) return TRIS_ERR_SHAPE;
    if (g->b_mode == TRIS_OP_CONV) {
        const int C = g->N;
        if (C % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: conv channels %d %% 8", C);
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.img_w, (uint64_t)p.img_h, (uint64_t)p.img_n};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.img_w * C * 2, (uint64_t)p.img_h * p.img_w * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
        mb = tris::tensor_map_bf16(g->b, 4, dims, str, box);
    } else if (g->b_mode == TRIS_OP_K2D) {
        if (g->ldb % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: ldb %d %% 8", g->ldb);
        uint64_t dims[2] = {(uint64_t)g->K, (uint64_t)g->N};
        uint64_t str[1] = {(uint64_t)g->ldb * 2};
        uint32_t box[2] = {64, (uint32_t)bn};
        mb = tris::tensor_map_bf16(g->b, 2, dims, str, box);
    } else {
        if (g->ldb % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: ldb %d %% 8", g->ldb);
        // [K rows, ldb] row-major; for a conv dgrad the contiguous dim holds taps x N.
        const uint64_t inner = (conv && !wgrad) ? (uint64_t)p.taps * g->b_tap_stride : (uint64_t)g->N;
        const uint64_t outer = (conv && !wgrad) ? (uint64_t)a_ch : (uint64_t)g->K;
        uint64_t dims[2] = {inner, outer};
        uint64_t str[1] = {(uint64_t)g->ldb * 2};
        uint32_t box[2] = {64, (uint32_t)p.bk};
        mb = tris::tensor_map_bf16(g->b, 2, dims, str, box);
    }
    if (!mb) return TRIS_ERR_SHAPE;

    static bool attr_set = false;
    if (!attr_set) {
        TRIS_CUDA_OK(cudaFuncSetAttribute(tris_umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int total_tiles = p.tiles_tap * p.tiles_m * p.tiles_n * p.split_k;
    int ctas = tris::sm_count();
    if (g->max_ctas > 0 && g->max_ctas < ctas) ctas = g->max_ctas;
    if (total_tiles < ctas) ctas = total_tiles;
    tris_umma_gemm_kernel<<<ctas, kThreads, smem_bytes, stream>>>(*ma, *mb, p);
    TRIS_LAUNCH_OK("tris_umma_gemm_kernel");
    return TRIS_OK;
}
