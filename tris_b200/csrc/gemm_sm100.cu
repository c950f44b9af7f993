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
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 352;   // warp 0 TMA(A), warp 1 MMA, warps 2-9 epilogue, warp 10 TMA(B)
constexpr int kMaxStages = 9;    // barrier slots; the streaming rings use at most kRingStages, the resident-weight form all nine
constexpr int kRingStages = 8;
constexpr int kTmemCols = 512;
constexpr int kMaxAcc = 4;       // TMEM accumulator ring: min(4, 512 / BN) buffers of BN fp32 columns

// Division by a launch-time constant without the ~25-instruction integer divide (these run in single-thread loops).
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ int div(int x) const { return d == 1 ? x : static_cast<int>(__umulhi(static_cast<uint32_t>(x), mul) >> shr); }
    __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = div(x); r = x - q * static_cast<int>(d); }
};
static FastDiv make_fastdiv(int d) {
    FastDiv f{static_cast<uint32_t>(d < 1 ? 1 : d), 0, 0};
    if (f.d > 1) {
        uint32_t lg = 0;
        while ((1u << lg) < f.d) ++lg;
        const uint32_t pw = 31 + lg;
        f.mul = static_cast<uint32_t>(((1ull << pw) + f.d - 1) / f.d);
        f.shr = pw - 32;
    }
    return f;
}

struct KParams {
    int M, N, K;
    int a_mode, b_mode, wgrad, flip, taps;
    int tiles_m, tiles_n, tiles_tap, split_k, kblocks, kb_per_split;
    int batch, a_batched, b_batched;
    int bk, n_mma, bn;
    int img_n, img_h, img_w, th, tw, tiles_h, tiles_w, cblocks, b_tap_stride;
    uint32_t a_bytes, b_bytes, a_atom, b_atom, a_tx, b_tx, stages, staging_bytes, stats_bytes;
    uint32_t nacc, acc_stride, nstg;
    FastDiv fd_tiles_n, fd_tiles_m, fd_tiles_tap, fd_split, fd_tiles_w, fd_tiles_h, fd_cblocks, fd_tw, fd_bn;
    long long* dbg_buf;   // dbg & 8: per-tile clock64 stamps of CTA 0 ([role][64 tiles])
    int dbg;   // TRIS_GEMM_DEBUG bit mask (profiling aid): 1 skip TMA loads, 2 skip MMA issue, 4 skip epilogue work   // accumulator ring size / column stride, staging buffers (1 or 2)
    uint32_t idesc;
    void* d;
    const float* bias;
    const __nv_bfloat16* residual;
    float* stats;                    // [stats_parts][2N] per-CTA partial column statistics (plain stores, fixed order)
    int stats_parts, stats_mode;     // mode 0: (sum x, sum x^2) of the stored tile; 1: (sum g, sum g*(y - mu)) (BatchNorm backward)
    const __nv_bfloat16* stats_y;    // mode 1: y [M, ldd] (conv mode: NHWC with ldd channels)
    const float* stats_mu;           // mode 1: per-column mean [N]
    const float* mask_sc;            // optional [N]: multiply the output by ((stats_y * mask_sc + mask_sh) > 0)  (ReLU mask
    const float* mask_sh;            //   of relu(bn(y)) recomputed from y)
    // LD_CONV_HALO: the (th+2) x (tw+2) pixel box of a tile is loaded ONCE per 64-channel block; the nine taps are nine start
    // addresses into it (K-major SWIZZLE_128B operands work with row-shifted starts and a non-1024 group stride: the swizzle is
    // a function of the absolute smem address -- tools/micro/umma_shift.cu)
    uint32_t a_stages, a_stage_bytes, halo_pitch;   // ring depth, bytes per halo tile (1024-aligned), tw + 2
    uint32_t ring_bytes;             // bytes of all operand rings (staging starts there)
    int ng;                          // epilogue warp groups (1 or 2): with 2, the two groups of four warps drain ALTERNATING tiles,
                                     // so the latency chain of one tile's epilogue overlaps the other's (short-K GEMMs)
    int b_stationary;                // halo form, 64 input channels, one N tile: the nine weight tiles are loaded ONCE per CTA
    int split_ws;                    // 1 = split-K partials leave through plain TMA stores into a [split][...] workspace
    const unsigned char* res_bits;   // optional [M, N/8]: residual added only where its bit is set
    // EPI 2 (tile-local InstanceNorm: batched mode, one tile = the M <= 128 pixels of one image x bn channels)
    const float* in_gamma; const float* in_beta;   // [N]
    float* in_mean; float* in_invstd;               // [batch, N] saved for the backward pass
    const __nv_bfloat16* in_add;                    // optional bf16 [batch*M, ldd]: added after the mix scale
    float in_eps, in_mix;
    int in_relu;
    unsigned long long* tstamp;      // optional [2]: min(%globaltimer at CTA start), max(%globaltimer at CTA end) of this launch
    __nv_bfloat16* d_pre;            // optional: pre-activation (post-bias) output, bf16 [M, ldd]
    const __nv_bfloat16* dact_src;   // optional: multiply by act'(dact_src[row, col]) instead of applying act
    int ldd, act, out_f32, atomic, res_f32;
    float scale;                     // accumulator scale applied before the bias (1 = none)
};

struct SmemCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_full[kMaxAcc];
    uint64_t acc_empty[kMaxAcc];
    uint64_t afull[4];    // LD_CONV_HALO: ring of halo tiles (A), decoupled from the per-tap weight ring (full / empty)
    uint64_t aempty[4];
    uint32_t tmem_base;
};

struct TileCoord {
    int m_t, n_t, tap, split, batch;
    int n_i, h0, w0;  // conv fwd/dgrad output patch
};

__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
    TileCoord c;
    p.fd_tiles_n.divmod(t, t, c.n_t);
    p.fd_tiles_m.divmod(t, t, c.m_t);
    p.fd_tiles_tap.divmod(t, t, c.tap);
    p.fd_split.divmod(t, c.batch, c.split);
    c.n_i = c.h0 = c.w0 = 0;
    if (p.a_mode == TRIS_OP_CONV && !p.wgrad) {
        int r, tw_i, th_i;
        p.fd_tiles_w.divmod(c.m_t, r, tw_i);
        p.fd_tiles_h.divmod(r, c.n_i, th_i);
        c.w0 = tw_i * p.tw;
        c.h0 = th_i * p.th;
    }
    return c;
}

__device__ __forceinline__ uint32_t raw_bits(float f) { return __float_as_uint(f); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
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

// Loader variants (compile-time, so the single-thread producer / issuer loops carry no mode branches or divisions)
enum { LD_K2D = 0, LD_MN2D = 1, LD_CONV = 2, LD_CONV_WG = 3, LD_MN2D_TAPS = 4, LD_CONV_HALO = 5 };

// EPI = 3: the residual is fp32 (residual stream of the transformer towers).
// EPI = 1 adds the BatchNorm-backward epilogue (recomputed ReLU mask from y, sums (g, g (y - mu))): compiled separately so
// that the common kernels do not carry its code.
template <int AM, int BM, int EPI>
// 96 registers/thread (no spills; 166 unconstrained): 352 x 96 = 33 K registers leave room for one 256-thread CTA of the
// HBM-bound BatchNorm kernels next to a resident GEMM CTA, so side-stream weight gradients and the BN chain share SMs.
#ifndef TRIS_GEMM_MAXNREG
#define TRIS_GEMM_MAXNREG 96
#endif
__global__ void __maxnreg__(TRIS_GEMM_MAXNREG)
tris_umma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_e, const KParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + p.ring_bytes + p.nstg * p.staging_bytes + p.stats_bytes);
    const uint32_t smem_base = ptx::smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_tap * p.tiles_m * p.tiles_n * p.split_k * p.batch;

    if (threadIdx.x == 0) {
        if (p.tstamp != nullptr) {   // measurement aid (bench.py): device-side start of this launch
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMin(p.tstamp, t);
        }
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_b);
        ptx::prefetch_tmap(&map_d);
        for (uint32_t s = 0; s < p.stages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&ctl->full[s]), AM == LD_CONV_HALO ? 1 : 2);   // halo form: only the B producer arms it
            ptx::mbar_init(ptx::smem_u32(&ctl->empty[s]), 1);
        }
        for (int s = 0; s < 4; ++s) {
            ptx::mbar_init(ptx::smem_u32(&ctl->afull[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&ctl->aempty[s]), 1);
        }
        for (int b = 0; b < kMaxAcc; ++b) {
            ptx::mbar_init(ptx::smem_u32(&ctl->acc_full[b]), 1);
            ptx::mbar_init(ptx::smem_u32(&ctl->acc_empty[b]), 8 / p.ng);   // the warps of ONE epilogue group release a buffer
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
    // a kernel launched behind this one with programmatic stream serialization (the BatchNorm finalize kernels) may be set
    // up now; it blocks in griddepcontrol.wait until this grid has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 || warp == 10) {
        // ------------------------------------------------------------------ TMA producers: warp 0 feeds A, warp 10 feeds B
        // (two single-thread issue loops in parallel; each arms the stage's full barrier with its own byte count)
        if (AM == LD_CONV_HALO) {
            if (ptx::elect_one()) {
                const uint32_t b_ring = smem_base + p.a_stages * p.a_stage_bytes;
                if (warp == 0) {      // A: one halo box per (tile, 64-channel block)
                    uint32_t as = 0, aph = 0;
                    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                        const TileCoord c = decode_tile(p, t);
                        for (int cb = 0; cb < p.cblocks; ++cb) {
                            ptx::mbar_wait(ptx::smem_u32(&ctl->aempty[as]), aph ^ 1);
                            const uint32_t bar = ptx::smem_u32(&ctl->afull[as]);
                            ptx::mbar_expect_tx(bar, p.a_tx);
                            ptx::tma_load_4d(smem_base + as * p.a_stage_bytes, &map_a, bar, cb * 64, c.w0 - 1, c.h0 - 1, c.n_i);
                            if (++as == p.a_stages) { as = 0; aph ^= 1; }
                        }
                    }
                } else if (p.b_stationary) {   // B: the nine weight tiles of the layer, once (stage = tap, never released)
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t bar = ptx::smem_u32(&ctl->full[tap]);
                        const uint32_t sb = b_ring + tap * p.b_bytes;
                        ptx::mbar_expect_tx(bar, p.b_tx);
                        if (BM == LD_K2D) {
                            ptx::tma_load_2d(sb, &map_b, bar, tap * 64, 0);
                        } else {
                            for (int j = 0; j < p.bn / 64; ++j)
                                ptx::tma_load_2d(sb + j * p.b_atom, &map_b, bar, tap * p.b_tap_stride + 64 * j, 0);
                        }
                    }
                } else {              // B: one weight tile per (tile, channel block, tap)
                    uint32_t stage = 0, phase = 0;
                    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                        const TileCoord c = decode_tile(p, t);
                        const int n0 = c.n_t * p.bn;
                        for (int cb = 0; cb < p.cblocks; ++cb)
                            for (int tap = 0; tap < 9; ++tap) {
                                ptx::mbar_wait(ptx::smem_u32(&ctl->empty[stage]), phase ^ 1);
                                const uint32_t bar = ptx::smem_u32(&ctl->full[stage]);
                                const uint32_t sb = b_ring + stage * p.b_bytes;
                                ptx::mbar_expect_tx(bar, p.b_tx);
                                if (BM == LD_K2D) {
                                    ptx::tma_load_2d(sb, &map_b, bar, (tap * p.cblocks + cb) * 64, n0);
                                } else {
                                    const int inner = tap * p.b_tap_stride + n0;
                                    for (int j = 0; j < p.bn / 64; ++j)
                                        ptx::tma_load_2d(sb + j * p.b_atom, &map_b, bar, inner + 64 * j, cb * 64);
                                }
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            }
                    }
                }
            }
        } else if (ptx::elect_one()) {
            const bool is_a = warp == 0;
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t);
                const int m0 = c.m_t * kBlockM, n0 = c.n_t * p.bn;
                const int kb0 = c.split * p.kb_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                if ((p.dbg & 8) && blockIdx.x == 0) { const int i_ = (t - blockIdx.x) / gridDim.x; if (i_ < 64) p.dbg_buf[(is_a ? 0 : 1) * 64 + i_] = clock64(); }
                // incremental k-block state (no divisions inside the loop)
                int tap = 0, cb = kb0, dr = 0, ds = 0, pn = 0, ph0 = 0, pw0 = 0;
                if (AM == LD_CONV) {
                    p.fd_cblocks.divmod(kb0, tap, cb);
                    if (p.taps == 9) { dr = tap / 3 - 1; ds = tap % 3 - 1; }
                } else if (AM == LD_CONV_WG) {
                    int tw_i, r, th_i;
                    p.fd_tiles_w.divmod(kb0, r, tw_i);
                    p.fd_tiles_h.divmod(r, pn, th_i);
                    pw0 = tw_i * p.tw;
                    ph0 = th_i * p.th;
                    if (p.taps == 9) { dr = c.tap / 3 - 1; ds = c.tap % 3 - 1; }
                }
                const int sgn = p.flip ? -1 : 1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&ctl->empty[stage]), phase ^ 1);
                    const uint32_t bar = ptx::smem_u32(&ctl->full[stage]);
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = sa + p.a_bytes;
                    if (p.dbg & 1) {
                        ptx::mbar_arrive(bar);
                    } else if (is_a) {
                        ptx::mbar_expect_tx(bar, p.a_tx);
                        if (AM == LD_K2D) {
                            if (p.a_batched) ptx::tma_load_3d(sa, &map_a, bar, kb * 64, m0, c.batch);
                            else ptx::tma_load_2d(sa, &map_a, bar, kb * 64, m0);
                        } else if (AM == LD_MN2D) {
                            if (p.a_batched) {
                                ptx::tma_load_3d(sa, &map_a, bar, m0, kb * p.bk, c.batch);
                                ptx::tma_load_3d(sa + p.a_atom, &map_a, bar, m0 + 64, kb * p.bk, c.batch);
                            } else {
                                ptx::tma_load_2d(sa, &map_a, bar, m0, kb * p.bk);
                                ptx::tma_load_2d(sa + p.a_atom, &map_a, bar, m0 + 64, kb * p.bk);
                            }
                        } else if (AM == LD_CONV) {
                            ptx::tma_load_4d(sa, &map_a, bar, cb * 64, c.w0 + sgn * ds, c.h0 + sgn * dr, c.n_i);
                        } else {
                            ptx::tma_load_4d(sa, &map_a, bar, m0, pw0, ph0, pn);
                            ptx::tma_load_4d(sa + p.a_atom, &map_a, bar, m0 + 64, pw0, ph0, pn);
                        }
                    } else {
                        ptx::mbar_expect_tx(bar, p.b_tx);
                        if (BM == LD_K2D) {
                            if (p.b_batched) ptx::tma_load_3d(sb, &map_b, bar, kb * 64, n0, c.batch);
                            else ptx::tma_load_2d(sb, &map_b, bar, kb * 64, n0);
                        } else if (BM == LD_MN2D) {
                            for (int j = 0; j < p.bn / 64; ++j) {
                                if (p.b_batched) ptx::tma_load_3d(sb + j * p.b_atom, &map_b, bar, n0 + 64 * j, kb * p.bk, c.batch);
                                else ptx::tma_load_2d(sb + j * p.b_atom, &map_b, bar, n0 + 64 * j, kb * p.bk);
                            }
                        } else if (BM == LD_MN2D_TAPS) {
                            const int inner = tap * p.b_tap_stride + n0;
                            for (int j = 0; j < p.bn / 64; ++j)
                                ptx::tma_load_2d(sb + j * p.b_atom, &map_b, bar, inner + 64 * j, cb * 64);
                        } else {
                            for (int j = 0; j < p.bn / 64; ++j)
                                ptx::tma_load_4d(sb + j * p.b_atom, &map_b, bar, n0 + 64 * j, pw0 + ds, ph0 + dr, pn);
                        }
                    }
                    if (AM == LD_CONV) {
                        if (++cb == p.cblocks) {
                            cb = 0;
                            ++tap;
                            if (++ds == 2) { ds = -1; ++dr; }
                        }
                    } else if (AM == LD_CONV_WG) {
                        pw0 += p.tw;
                        if (pw0 >= p.tiles_w * p.tw) {
                            pw0 = 0;
                            ph0 += p.th;
                            if (ph0 >= p.tiles_h * p.th) { ph0 = 0; ++pn; }
                        }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer (single elected thread)
        if (AM == LD_CONV_HALO) {
            if (ptx::elect_one()) {
                uint32_t stage = 0, phase = 0, as = 0, aph = 0, acc_phase = 0;
                int it = 0;
                constexpr bool b_mn = (BM != LD_K2D);
                constexpr uint32_t b_kstep = (b_mn ? 2048u : 32u) >> 4;
                const uint64_t da0 = ptx::umma_smem_desc_sw128(0, 0u, p.halo_pitch * 128);   // 8-row groups = patch rows, pitch tw + 2
                const uint64_t db0 = ptx::umma_smem_desc_sw128(0, b_mn ? p.b_atom : 0u, 1024);
                const uint32_t b_ring = smem_base + p.a_stages * p.a_stage_bytes;
                const int sgn = p.flip ? -1 : 1;
                for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                    const int buf = it & (p.nacc - 1);
                    ptx::mbar_wait(ptx::smem_u32(&ctl->acc_empty[buf]), ((acc_phase >> buf) & 1) ^ 1);
                    acc_phase ^= 1u << buf;
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * p.acc_stride;
                    uint32_t accum = 0;
                    for (int cb = 0; cb < p.cblocks; ++cb) {
                        ptx::mbar_wait(ptx::smem_u32(&ctl->afull[as]), aph);
                        const uint32_t a_tile = smem_base + as * p.a_stage_bytes;
                        if (p.b_stationary) {
                            // resident weights (stage = tap): their barriers completed phase 0 once and for all
                            if (it == 0) {
                                for (int tap = 0; tap < 9; ++tap) ptx::mbar_wait(ptx::smem_u32(&ctl->full[tap]), 0);
                            }
                            ptx::tc_fence_after();
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const int dr = tap / 3 - 1, ds = tap % 3 - 1;
                                const uint32_t sa = (a_tile + static_cast<uint32_t>((1 + sgn * dr) * static_cast<int>(p.halo_pitch) + 1 + sgn * ds) * 128u) >> 4;
                                const uint32_t sb = (b_ring + tap * p.b_bytes) >> 4;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    ptx::umma_f16(tmem_d, da0 | static_cast<uint64_t>(sa + k * 2), db0 | static_cast<uint64_t>(sb + k * b_kstep), p.idesc, accum);
                                    accum = 1;
                                }
                            }
                        } else {
                            for (int tap = 0; tap < 9; ++tap) {
                                const int dr = tap / 3 - 1, ds = tap % 3 - 1;
                                const uint32_t sa = (a_tile + static_cast<uint32_t>((1 + sgn * dr) * static_cast<int>(p.halo_pitch) + 1 + sgn * ds) * 128u) >> 4;
                                ptx::mbar_wait(ptx::smem_u32(&ctl->full[stage]), phase);
                                ptx::tc_fence_after();
                                const uint32_t sb = (b_ring + stage * p.b_bytes) >> 4;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    ptx::umma_f16(tmem_d, da0 | static_cast<uint64_t>(sa + k * 2), db0 | static_cast<uint64_t>(sb + k * b_kstep), p.idesc, accum);
                                    accum = 1;
                                }
                                ptx::umma_commit(ptx::smem_u32(&ctl->empty[stage]));
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            }
                        }
                        ptx::umma_commit(ptx::smem_u32(&ctl->aempty[as]));
                        if (++as == p.a_stages) { as = 0; aph ^= 1; }
                    }
                    ptx::umma_commit(ptx::smem_u32(&ctl->acc_full[buf]));
                }
            }
        } else if (ptx::elect_one()) {
            uint32_t stage = 0, phase = 0;
            uint32_t acc_phase = 0;   // bit b = phase of accumulator buffer b
            int it = 0;
            constexpr bool a_mn = (AM == LD_MN2D) || (AM == LD_CONV_WG);
            constexpr bool b_mn = (BM != LD_K2D);
            constexpr uint32_t a_kstep = (a_mn ? 2048u : 32u) >> 4, b_kstep = (b_mn ? 2048u : 32u) >> 4;
            // descriptor templates: only the 14-bit start-address field changes per stage / k-step
            const uint64_t da0 = ptx::umma_smem_desc_sw128(0, a_mn ? p.a_atom : 0u, 1024);
            const uint64_t db0 = ptx::umma_smem_desc_sw128(0, b_mn ? p.b_atom : 0u, 1024);
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const TileCoord c = decode_tile(p, t);
                const int kb0 = c.split * p.kb_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                const int buf = it & (p.nacc - 1);
                if ((p.dbg & 8) && blockIdx.x == 0 && it < 64) p.dbg_buf[2 * 64 + it] = clock64();
                ptx::mbar_wait(ptx::smem_u32(&ctl->acc_empty[buf]), ((acc_phase >> buf) & 1) ^ 1);
                if ((p.dbg & 8) && blockIdx.x == 0 && it < 64) p.dbg_buf[3 * 64 + it] = clock64();
                acc_phase ^= 1u << buf;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * p.acc_stride;
                uint32_t accum = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&ctl->full[stage]), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = (smem_base + stage * stage_bytes) >> 4;
                    const uint32_t sb = sa + (p.a_bytes >> 4);
#pragma unroll 4
                    for (int k = 0; k < ((p.dbg & 2) ? 0 : p.n_mma); ++k) {
                        ptx::umma_f16(tmem_d, da0 | static_cast<uint64_t>(sa + k * a_kstep), db0 | static_cast<uint64_t>(sb + k * b_kstep),
                                      p.idesc, accum);
                        accum = 1;
                    }
                    if (p.dbg & 16) ptx::mbar_arrive(ptx::smem_u32(&ctl->empty[stage]));
                    else ptx::umma_commit(ptx::smem_u32(&ctl->empty[stage]));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(ptx::smem_u32(&ctl->acc_full[buf]));
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ------------------------------------------------------------------ epilogue: 8 warps.  Warp w may touch TMEM
        // lanes 32*(w%4)..+31; the two warps of a lane quadrant split the 32-column chunks (even / odd).  Values are
        // staged in 128B-swizzled smem ([group of 128 B columns][128 rows]) and leave through TMA stores (coalesced,
        // edge-clipped; fp32 weight gradients use TMA reduce-add instead of atomics).
        const int ew = warp - 2;
        const int q = warp & 3;
        const int ng = p.ng;                              // 1: all eight warps per tile; 2: four warps per tile, alternating tiles
        const int GT = 256 / ng;                          // threads per group
        const int grp = ng == 2 ? (ew >> 2) : 0;
        const int half = ng == 2 ? 0 : (ew >> 2);         // one group: the two warps of a lane quadrant split the chunks
        const int chstep = ng == 2 ? 1 : 2;
        const int tid_e = ng == 2 ? ((threadIdx.x - 64) & 127) : (threadIdx.x - 64);   // thread index inside the group
        const bool lead = (ng == 2 ? (ew & 3) : ew) == 0; // first warp of the group (its elected lane owns the TMA stores)
        const uint32_t bar_id = 1 + grp;
        const int row = q * 32 + lane;
        const uint32_t stg0 = smem_base + p.ring_bytes;
        // per group: [2N] running column sums, then the scratch of the statistics pass
        float* s_stats = reinterpret_cast<float*>(smem + p.ring_bytes + p.nstg * p.staging_bytes) + grp * (p.stats_bytes / (4 * ng));
        const int esz = p.out_f32 ? 4 : 2;
        const int gw = 128 / esz;                      // columns per 128-byte staging group
        const int ngroups = (p.bn * esz) / 128;
        if (p.stats != nullptr) {
            for (int i = tid_e; i < 2 * p.N; i += GT) s_stats[i] = 0.f;
        }
        // statistics mapping (tile-invariant): thread = (16-byte vector of 8 columns, row part)
        const int nvec = p.bn >> 3;
        const int vec = tid_e % nvec, part = tid_e / nvec, nparts = GT / nvec;
        const int rpp = kBlockM / nparts;
        uint32_t acc_phase = 0;
        int it = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < total_tiles; t += ng * gridDim.x, it += ng) {
            const TileCoord c = decode_tile(p, t);
            const int buf = it & (p.nacc - 1);
            const uint32_t stg = stg0 + (p.nstg == 2 ? (it & 1) * p.staging_bytes : 0);
            const int n0 = c.n_t * p.bn;
            const long grow = static_cast<long>(c.m_t) * kBlockM + row;   // plain-mode row (residual / d_pre / dact)
            const bool rvalid = grow < p.M;
            // One thread polls the mbarrier (256 pollers would saturate the SM's barrier unit and slow the producer /
            // issuer threads); the named barrier then releases the other epilogue warps.
            // (elected lane of the first epilogue warp: elect.sync always picks the same lane, which therefore also owns the
            // bulk async-groups of the TMA stores below)
            if (lead && ptx::elect_one()) {
                if ((p.dbg & 8) && blockIdx.x == 0 && it < 64) p.dbg_buf[4 * 64 + it] = clock64();
                ptx::mbar_wait(ptx::smem_u32(&ctl->acc_full[buf]), (acc_phase >> buf) & 1);
                if ((p.dbg & 8) && blockIdx.x == 0 && it < 64) p.dbg_buf[5 * 64 + it] = clock64();
                // the TMA stores that last used this staging buffer have finished reading it
                // (two groups: each owns one staging buffer and its own bulk groups)
                if (p.nstg == 2 && ng == 1) ptx::bulk_wait_read1(); else ptx::bulk_wait_read0();
            }
            acc_phase ^= 1u << buf;
            ptx::named_bar_sync(bar_id, GT);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * p.acc_stride;
            for (int ch = half; ch < ((p.dbg & 4) ? 0 : p.bn / 32); ch += chstep) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(taddr + ch * 32, raw);
                ptx::tmem_ld_wait();
                const int ncol0 = n0 + ch * 32;  // column within N
                if (ncol0 >= p.N) continue;      // warp-uniform
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * p.scale;
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ncol0 + i < p.N) v[i] += __ldg(p.bias + ncol0 + i);
                }
                const long off = grow * p.ldd + ncol0;
                if (EPI == 1 && p.mask_sc != nullptr) {
                    // g = acc * [relu(bn(y)) > 0], the mask recomputed from y (one 64-byte row segment per thread)
                    long yrow = grow;
                    bool yok = rvalid;
                    if (AM == LD_CONV || AM == LD_CONV_HALO) {
                        int qh, qw;
                        p.fd_tw.divmod(row, qh, qw);
                        yok = row < p.th * p.tw && c.h0 + qh < p.img_h && c.w0 + qw < p.img_w;
                        yrow = (static_cast<long>(c.n_i) * p.img_h + c.h0 + qh) * p.img_w + c.w0 + qw;
                    }
                    if (yok) {
                        const uint4* yp = reinterpret_cast<const uint4*>(p.stats_y + yrow * p.ldd + ncol0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (ncol0 + g * 8 < p.N) {
                                uint4 rr = __ldg(yp + g);
                                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 f = __bfloat1622float2(r2[j]);
                                    const int cc = ncol0 + g * 8 + 2 * j;
                                    if (fmaf(f.x, __ldg(p.mask_sc + cc), __ldg(p.mask_sh + cc)) <= 0.f) v[g * 8 + 2 * j] = 0.f;
                                    if (fmaf(f.y, __ldg(p.mask_sc + cc + 1), __ldg(p.mask_sh + cc + 1)) <= 0.f) v[g * 8 + 2 * j + 1] = 0.f;
                                }
                            }
                        }
                    }
                }
                if (p.d_pre != nullptr && rvalid) {
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
                auto add_residual = [&]() {
                    if (EPI == 3) {      // fp32 residual (instantiated separately: the hot instantiations keep their code)
                      if (p.residual != nullptr && rvalid) {
                        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + off);
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            if (ncol0 + g * 4 < p.N) {
                                const float4 f = __ldg(rp + g);
                                v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
                            }
                        }
                      }
                    } else if (p.residual != nullptr && rvalid) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
                        uint32_t bits = 0xffffffffu;
                        if (p.res_bits != nullptr) bits = __ldg(reinterpret_cast<const uint32_t*>(p.res_bits + grow * (p.N >> 3) + (ncol0 >> 3)));
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (ncol0 + g * 8 < p.N) {
                                uint4 rr = __ldg(rp + g);
                                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 f = __bfloat1622float2(r2[j]);
                                    v[g * 8 + 2 * j] += ((bits >> (g * 8 + 2 * j)) & 1u) ? f.x : 0.f;
                                    v[g * 8 + 2 * j + 1] += ((bits >> (g * 8 + 2 * j + 1)) & 1u) ? f.y : 0.f;
                                }
                            }
                        }
                    }
                };
                if (p.dact_src != nullptr) {
                    // derivative form: d = (acc + residual) * act'(dact_src) -- the residual (the other branch of a join) is
                    // part of the gradient that passes through the activation
                    add_residual();
                    if (rvalid) {
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
                    }
                } else if (p.act != TRIS_ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = act_apply(v[i], p.act);
                }
                if (p.dact_src == nullptr) add_residual();
                // ---- stage (swizzle: 16-byte chunk index ^ (row & 7), the TMA SWIZZLE_128B pattern)
                if (p.out_f32) {
                    const uint32_t base = stg + ch * 16384 + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        ptx::st_shared_v4(base + ((j ^ (row & 7)) << 4), raw_bits(v[4 * j]), raw_bits(v[4 * j + 1]),
                                          raw_bits(v[4 * j + 2]), raw_bits(v[4 * j + 3]));
                } else {
                    const uint32_t base = stg + (ch >> 1) * 16384 + row * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = (ch & 1) * 4 + j;
                        ptx::st_shared_v4(base + ((idx ^ (row & 7)) << 4), pack_bf16(v[8 * j], v[8 * j + 1]),
                                          pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                          pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    }
                }
            }
            // accumulator buffer is free for the MMA warp as soon as every epilogue warp has read its part
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->acc_empty[buf]));
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(bar_id, GT);
            // ---- TMA store (one elected thread), issued BEFORE the statistics pass: both only read the staged tile, so the
            // store drains while the epilogue warps accumulate the column sums
            if (lead && !(p.dbg & (4 | 64)) && ptx::elect_one()) {
                for (int g = 0; g < ngroups; ++g) {
                    const int cg = n0 + g * gw;
                    if (cg >= p.N) break;
                    const uint32_t src = stg + g * 16384;
                    if (AM == LD_CONV || AM == LD_CONV_HALO) {
                        ptx::tma_store_4d(&map_d, src, cg, c.w0, c.h0, c.n_i);
                    } else if (p.wgrad) {
                        if (p.split_ws) ptx::tma_store_4d(&map_d, src, cg, c.tap, c.m_t * kBlockM, c.split);
                        else ptx::tma_reduce_add_3d(&map_d, src, cg, c.tap, c.m_t * kBlockM);
                    } else if (p.split_ws) {
                        ptx::tma_store_3d(&map_d, src, cg, c.m_t * kBlockM, c.split);
                    } else if (p.batch > 1) {
                        if (p.atomic) ptx::tma_reduce_add_3d(&map_d, src, cg, c.m_t * kBlockM, c.batch);
                        else ptx::tma_store_3d(&map_d, src, cg, c.m_t * kBlockM, c.batch);
                    } else if (p.atomic) {
                        ptx::tma_reduce_add_2d(&map_d, src, cg, c.m_t * kBlockM);
                    } else {
                        ptx::tma_store_2d(&map_d, src, cg, c.m_t * kBlockM);
                    }
                }
                ptx::bulk_commit();
            }
            // ---- BatchNorm column statistics of the stored (bf16-rounded) tile.  Thread = (16-byte vector of 8 columns,
            // row part): one LDS.128 + 3 instructions per element; parts are combined through a 16 KB scratch and the
            // per-CTA totals live in smem until the kernel ends (one global atomic per column per CTA).
            if ((p.stats != nullptr || EPI == 2) && !(p.dbg & (4 | 32))) {
                int rlim = kBlockM;
                constexpr bool conv_tile = (AM == LD_CONV) || (AM == LD_CONV_HALO);
                if (conv_tile) rlim = p.th * p.tw;
                else rlim = min(kBlockM, p.M - c.m_t * kBlockM);
                float sa[8], sq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) sa[i] = sq[i] = 0.f;
                const int r0 = part * rpp, r1 = min(rlim, r0 + rpp);
                int hh = 0, ww = 0;
                if (conv_tile) { int qh, qw; p.fd_tw.divmod(r0, qh, qw); hh = c.h0 + qh; ww = c.w0 + qw; }
                const uint32_t gbase = stg + (vec >> 3) * 16384;
                const int idx = vec & 7;
                const int scol = n0 + vec * 8;
                float mu[8];
                if (EPI == 1 && p.stats_mode == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) mu[i] = scol + i < p.N ? __ldg(p.stats_mu + scol + i) : 0.f;
                }
                for (int r = r0; r < r1; ++r) {
                    // patch rows that overhang the image are computed (halo taps see real pixels) but never stored
                    const bool ok = !conv_tile || (hh < p.img_h && ww < p.img_w);
                    const long yrow = conv_tile ? (static_cast<long>(c.n_i) * p.img_h + hh) * p.img_w + ww
                                                : static_cast<long>(c.m_t) * kBlockM + r;
                    if (conv_tile && ++ww == c.w0 + p.tw) { ww = c.w0; ++hh; }
                    if (!ok) continue;
                    uint32_t w4[4];
                    ptx::ld_shared_v4(gbase + r * 128 + ((idx ^ (r & 7)) << 4), w4);
                    if (EPI == 1 && p.stats_mode == 1) {
                        if (scol < p.N) {
                            const uint4 yy = __ldg(reinterpret_cast<const uint4*>(p.stats_y + yrow * p.ldd + scol));
                            const uint32_t y4[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float x0 = __uint_as_float(w4[k] << 16), x1 = __uint_as_float(w4[k] & 0xffff0000u);
                                const float y0 = __uint_as_float(y4[k] << 16), y1 = __uint_as_float(y4[k] & 0xffff0000u);
                                sa[2 * k] += x0; sq[2 * k] = fmaf(x0, y0 - mu[2 * k], sq[2 * k]);
                                sa[2 * k + 1] += x1; sq[2 * k + 1] = fmaf(x1, y1 - mu[2 * k + 1], sq[2 * k + 1]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float x0 = __uint_as_float(w4[k] << 16), x1 = __uint_as_float(w4[k] & 0xffff0000u);
                            sa[2 * k] += x0; sq[2 * k] = fmaf(x0, x0, sq[2 * k]);
                            sa[2 * k + 1] += x1; sq[2 * k + 1] = fmaf(x1, x1, sq[2 * k + 1]);
                        }
                    }
                }
                float* scr = s_stats + (EPI == 2 ? 2 * p.bn : 2 * p.N);   // [nparts][2][bn] (inside this group's region)
                float* dst = scr + (part * 2) * p.bn + vec * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) { dst[i] = sa[i]; dst[p.bn + i] = sq[i]; }
                ptx::named_bar_sync(bar_id, GT);
                if (EPI != 2) {
                    for (int jx = tid_e; jx < 2 * p.bn; jx += GT) {
                        int which, col;
                        p.fd_bn.divmod(jx, which, col);
                        if (n0 + col < p.N) {
                            float tot = 0.f;
                            for (int pp = 0; pp < nparts; ++pp) tot += scr[(pp * 2 + which) * p.bn + col];
                            s_stats[which * p.N + n0 + col] += tot;      // single owner per column within the CTA
                        }
                    }
                } else {
                    // ---- tile-local InstanceNorm (model/attn.py:75-105): this tile holds ALL pixels of one image for bn
                    // channels, so mean / biased variance over the pixels are complete here.  s_stats[0..bn) = scale
                    // (gamma * invstd), s_stats[bn..2bn) = shift (beta - mean * scale); mean / invstd saved for backward.
                    for (int col = tid_e; col < p.bn; col += GT) {
                        if (n0 + col < p.N) {
                            float sum = 0.f, sq = 0.f;
                            for (int pp = 0; pp < nparts; ++pp) { sum += scr[(pp * 2) * p.bn + col]; sq += scr[(pp * 2 + 1) * p.bn + col]; }
                            const float inv_n = 1.f / static_cast<float>(rlim);
                            const float mean = sum * inv_n;
                            const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
                            const float invstd = rsqrtf(var + p.in_eps);
                            const float sc = __ldg(p.in_gamma + n0 + col) * invstd;
                            s_stats[col] = sc;
                            s_stats[p.bn + col] = __ldg(p.in_beta + n0 + col) - mean * sc;
                            p.in_mean[static_cast<long>(c.batch) * p.N + n0 + col] = mean;
                            p.in_invstd[static_cast<long>(c.batch) * p.N + n0 + col] = invstd;
                        }
                    }
                    // the raw tile's TMA store must have finished READING the staging buffer before it is normalised in place
                    if (lead && ptx::elect_one()) ptx::bulk_wait_read0();
                    ptx::named_bar_sync(bar_id, GT);
                    if (scol < p.N) {
                        float sc8[8], sh8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { sc8[i] = s_stats[vec * 8 + i]; sh8[i] = s_stats[p.bn + vec * 8 + i]; }
                        for (int r = r0; r < r1; ++r) {
                            const uint32_t addr = gbase + r * 128 + ((idx ^ (r & 7)) << 4);
                            uint32_t w4[4];
                            ptx::ld_shared_v4(addr, w4);
                            float o[8];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                o[2 * k] = fmaf(__uint_as_float(w4[k] << 16), sc8[2 * k], sh8[2 * k]);
                                o[2 * k + 1] = fmaf(__uint_as_float(w4[k] & 0xffff0000u), sc8[2 * k + 1], sh8[2 * k + 1]);
                            }
                            if (p.in_relu) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) o[i] *= p.in_mix;
                            if (p.in_add != nullptr) {
                                const long arow = static_cast<long>(c.batch) * p.M + r;
                                const uint4 aa = __ldg(reinterpret_cast<const uint4*>(p.in_add + arow * p.ldd + scol));
                                const uint32_t a4[4] = {aa.x, aa.y, aa.z, aa.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    o[2 * k] += __uint_as_float(a4[k] << 16);
                                    o[2 * k + 1] += __uint_as_float(a4[k] & 0xffff0000u);
                                }
                            }
                            ptx::st_shared_v4(addr, pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                        }
                    }
                    ptx::fence_proxy_async_smem();
                    ptx::named_bar_sync(bar_id, GT);
                    if (lead && ptx::elect_one()) {
                        for (int g = 0; g < ngroups; ++g) {
                            const int cg = n0 + g * gw;
                            if (cg >= p.N) break;
                            ptx::tma_store_3d(&map_e, stg + g * 16384, cg, c.m_t * kBlockM, c.batch);
                        }
                        ptx::bulk_commit();
                    }
                }
            }
            if ((p.dbg & 8) && blockIdx.x == 0 && tid_e == 0 && it < 64) p.dbg_buf[6 * 64 + it] = clock64();
        }
        if (lead && ptx::elect_one()) ptx::bulk_wait_all();
        if (p.stats != nullptr) {
            // deterministic: every CTA stores its partial sums as row blockIdx.x of stats[stats_parts][2N] (tiles are
            // assigned statically, sums inside a CTA run in a fixed order); the consumer adds the rows in order.
            // Rows no CTA owns (grid smaller than stats_parts) are cleared here.
            ptx::named_bar_sync(bar_id, GT);
            float* dst = p.stats + static_cast<long>(blockIdx.x) * 2 * p.N;
            if (ng == 2) {
                // the two groups accumulated separately: meet on a CTA-wide named barrier, group 0 adds (fixed order)
                ptx::named_bar_sync(3, 256);
                const float* other = s_stats + p.stats_bytes / 8;
                if (grp == 0)
                    for (int i = tid_e; i < 2 * p.N; i += GT) dst[i] = s_stats[i] + other[i];
            } else {
                for (int i = tid_e; i < 2 * p.N; i += GT) dst[i] = s_stats[i];
            }
            if (grp == 0)
                for (int r = gridDim.x + blockIdx.x; r < p.stats_parts; r += gridDim.x) {
                    float* z = p.stats + static_cast<long>(r) * 2 * p.N;
                    for (int i = tid_e; i < 2 * p.N; i += GT) z[i] = 0.f;
                }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
    if (p.tstamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(p.tstamp + 1, t);
    }
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Second stage of a split-K weight gradient: d[m, j] (+)= sum_s ws[s][m][j], the partials added in split order (deterministic;
// replaces the arrival-order TMA reduce-adds).  ws rows are compact (`w` floats), d rows have stride ldd.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ d, int rows, int w,
                                                             int ldd, int split, int accumulate) {
    const long per = static_cast<long>(rows) * w;
    const long total4 = per >> 2;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long e = i << 2;
        const int m = static_cast<int>(e / w), j = static_cast<int>(e - static_cast<long>(m) * w);
        float4* dp = reinterpret_cast<float4*>(d + static_cast<long>(m) * ldd + j);
        float4 acc = accumulate ? *dp : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp0 = 0; sp0 < split; sp0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = sp0 + u < split ? __ldg(reinterpret_cast<const float4*>(ws + (sp0 + u) * per + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        *dp = acc;
    }
}

// Multi-tensor form: one launch reduces up to TRIS_REDUCE_MAX pending split-K weight gradients (blockIdx.y = item).
struct ReduceTable {
    tris_reduce_item it[TRIS_REDUCE_MAX];
};
__global__ void __launch_bounds__(256) splitk_reduce_multi_kernel(const ReduceTable t) {
    const tris_reduce_item& q = t.it[blockIdx.y];
    const float* __restrict__ ws = q.ws;
    float* __restrict__ d = q.d;
    const long per = static_cast<long>(q.rows) * q.w;
    const long total4 = per >> 2;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long e = i << 2;
        const int m = static_cast<int>(e / q.w), j = static_cast<int>(e - static_cast<long>(m) * q.w);
        float4* dp = reinterpret_cast<float4*>(d + static_cast<long>(m) * q.ldd + j);
        float4 acc = q.accumulate ? *dp : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp0 = 0; sp0 < q.split; sp0 += 8) {       // eight partials in flight, added in split order
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = sp0 + u < q.split ? __ldg(reinterpret_cast<const float4*>(ws + (sp0 + u) * per + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        *dp = acc;
    }
}

}  // namespace

extern "C" int tris_splitk_reduce_multi(const tris_reduce_item* items, int n, tris_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n < 0 || (n > 0 && !items)) return tris::fail(TRIS_ERR_SHAPE, "tris_splitk_reduce_multi: bad item list");
    for (int o = 0; o < n; o += TRIS_REDUCE_MAX) {
        ReduceTable t{};
        const int cnt = n - o < TRIS_REDUCE_MAX ? n - o : TRIS_REDUCE_MAX;
        long max4 = 0;
        for (int i = 0; i < cnt; ++i) {
            t.it[i] = items[o + i];
            if (!t.it[i].ws || !t.it[i].d || t.it[i].w % 4 || t.it[i].ldd % 4 || t.it[i].rows <= 0 || t.it[i].split < 1)
                return tris::fail(TRIS_ERR_SHAPE, "tris_splitk_reduce_multi: item %d malformed", o + i);
            const long n4 = static_cast<long>(t.it[i].rows) * t.it[i].w / 4;
            if (n4 > max4) max4 = n4;
        }
        // CTAs per item sized for the largest item (4 output vectors per thread); the CTAs of small items exit at once.
        // (32 CTAs per item left a 1024 x 1024 weight gradient to 8192 threads: 0.6 TB/s.)
        long gx = (max4 + 1023) / 1024;
        gx = gx < 32 ? 32 : (gx > 512 ? 512 : gx);
        splitk_reduce_multi_kernel<<<dim3(static_cast<unsigned>(gx), cnt), 256, 0, stream>>>(t);
        TRIS_LAUNCH_OK("splitk_reduce_multi_kernel");
    }
    return TRIS_OK;
}

extern "C" int tris_gemm(tris_gemm_desc* g, tris_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!g || !g->a || !g->b || !g->d) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: null operand");
    if (g->M <= 0 || g->N <= 0 || g->K <= 0) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: empty extent M=%d N=%d K=%d", g->M, g->N, g->K);
    const int bn = g->block_n;
    if (bn < 32 || bn > 256 || bn % 32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: block_n %d not in {32..256 step 32}", bn);
    // the 16-byte epilogue vectors (bf16 stores, residual / d_pre / dact_src loads) need N % 8; plain fp32 output does not
    const bool vec_epi = g->out_dtype != TRIS_DT_F32 || g->residual || g->d_pre || g->dact_src;
    if ((vec_epi && g->N % 8) || g->ldd % 4) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: N=%d (%%8) / ldd=%d (%%4) alignment", g->N, g->ldd);
    const bool conv = g->a_mode == TRIS_OP_CONV;
    const bool wgrad = conv && g->wgrad;
    const bool b_mn = g->b_mode != TRIS_OP_K2D;
    if (b_mn && bn % 64) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: MN-major B needs block_n %% 64 == 0");
    if ((g->b_mode == TRIS_OP_CONV) != wgrad) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: B conv mode only for wgrad");
    if (g->atomic && g->out_dtype != TRIS_DT_F32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: atomic needs f32 output");
    if (g->split_k > 1 && (g->out_dtype != TRIS_DT_F32 || !g->splitk_ws))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: split_k needs f32 output and a splitk_ws workspace");
    if (g->residual && g->out_dtype == TRIS_DT_F32 && g->atomic) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: residual+atomic");

    const int batch = g->batch > 1 ? g->batch : 1;
    if (batch > 1 && (conv || g->residual || g->d_pre || g->dact_src || g->stats))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: batched form supports the 2-D modes with bias/activation only");
    const bool inorm = g->d_norm != nullptr;
    if (inorm && (batch < 2 || g->a_mode != TRIS_OP_K2D || g->b_mode != TRIS_OP_K2D || g->M > kBlockM || g->out_dtype != TRIS_DT_BF16 ||
                  !g->in_gamma || !g->in_beta || !g->in_mean || !g->in_invstd || g->split_k > 1 || g->atomic || g->b_batch_stride != 0))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: the InstanceNorm epilogue needs the batched K-major form with M <= 128 rows per image, "
                                          "a shared B, bf16 output and gamma / beta / mean / invstd buffers");
    KParams p{};
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("TRIS_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
        p.dbg = dbg;
        p.dbg_buf = nullptr;
    }
    p.batch = batch;
    p.a_batched = batch > 1 && g->a_batch_stride != 0;
    p.b_batched = batch > 1 && g->b_batch_stride != 0;
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
    p.split_ws = p.split_k > 1 ? 1 : 0;
    g->split_used = p.split_k;
    if (p.split_ws && batch > 1) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: split_k with batch > 1 is not supported");
    // compact row width of the split-K workspace: [split][M][ws_w] fp32
    const int ws_w = wgrad ? p.taps * g->N : (g->N + 3) / 4 * 4;
    if (p.split_ws && (g->N % 4 || g->ldd % 4)) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: split_k needs N, ldd %% 4 == 0");

    const bool a_mn = (g->a_mode == TRIS_OP_MN2D) || wgrad;
    p.a_atom = p.bk * 128;
    p.b_atom = p.bk * 128;
    p.a_bytes = a_mn ? 2 * p.a_atom : kBlockM * 128;
    p.b_bytes = b_mn ? (bn / 64) * p.b_atom : bn * 128;
    // bytes actually delivered per stage (full boxes, OOB elements are zero-filled and counted)
    uint32_t a_tx = p.a_bytes, b_tx = p.b_bytes;
    if (conv && !wgrad) a_tx = p.th * p.tw * 128;
    const bool halo = conv && !wgrad && g->conv_halo != 0;
    if (halo) {
        if (p.taps != 9 || p.tw != 8 || p.th * p.tw != kBlockM)
            return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: the halo form needs a 3x3 conv and a 16 x 8 pixel tile (got %d x %d)", p.th, p.tw);
        p.halo_pitch = p.tw + 2;
        a_tx = (p.th + 2) * (p.tw + 2) * 128;
        p.a_stage_bytes = (a_tx + 1023) / 1024 * 1024;
        p.a_stages = 3;
        p.a_bytes = 0;           // the operand ring below holds the per-tap weight tiles only
    }
    p.a_tx = a_tx;
    p.b_tx = b_tx;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    const uint32_t halo_bytes = halo ? p.a_stages * p.a_stage_bytes : 0;
    const bool out_f32 = g->out_dtype == TRIS_DT_F32;
    if (out_f32 && bn > 128) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: fp32 output needs block_n <= 128 (staging)");
    if (conv && !wgrad && out_f32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: conv forward/dgrad output is bf16");
    if (conv && (g->residual || g->d_pre || g->dact_src || g->bias))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: bias/residual/d_pre/dact epilogues are for the 2-D modes");
    if (wgrad && (g->stats || g->mask_sc)) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: no stats / mask epilogue on weight gradients");
    if (wgrad && !out_f32) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: weight gradients are fp32");
    if (g->stats && (out_f32 || g->N > 2048)) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: stats need bf16 output, N <= 2048");
    if (g->stats && g->stats_parts < 1) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: stats need stats_parts >= 1 rows");
    if (g->stats_mode == 1 && (!g->stats || !g->stats_y || !g->stats_mu))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: stats_mode 1 needs stats, stats_y and stats_mu");
    if ((g->mask_sc != nullptr) != (g->mask_sh != nullptr) || (g->mask_sc && (!g->stats_y || out_f32)))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: the recomputed ReLU mask needs mask_sc, mask_sh, stats_y and bf16 output");
    if (g->mask_sc && batch > 1) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: mask epilogue is for the non-batched modes");
    p.staging_bytes = bn * (out_f32 ? 4 : 2) * 128;
    if (p.staging_bytes < 16384) p.staging_bytes = 16384;
    // column statistics: [2N] running sums + 16 KB scratch; with N <= 1024 room for a second copy of the sums, so that two
    // epilogue groups (p.ng == 2) can each own [2N sums][8 KB scratch]
    p.stats_bytes = g->stats ? ((2 * g->N * 4 + 1023) / 1024) * 1024 * (g->N <= 1024 ? 2 : 1) + 16384 : 0;
    if (inorm) p.stats_bytes = ((2 * bn * 4 + 1023) / 1024) * 1024 + 16384;
    p.nacc = 512 / bn >= 4 ? 4 : 2;   // power of two (ring index = it & (nacc - 1))
    p.acc_stride = bn;
    static uint32_t smem_cap = 0;   // TRIS_GEMM_SMEM_KB: leave shared memory for co-resident streaming kernels (experiment knob)
    if (smem_cap == 0) { const char* e = getenv("TRIS_GEMM_SMEM_KB"); smem_cap = (e && atoi(e) >= 64 && atoi(e) <= 227) ? atoi(e) * 1024u : 227u * 1024u; }
    const uint32_t fixed = 1024 + sizeof(SmemCtl) + 64 + p.stats_bytes;
    // two staging buffers (store of tile i overlaps the epilogue of tile i+1) when >= 4 pipeline stages still fit
    p.nstg = (smem_cap - fixed - halo_bytes - 2 * p.staging_bytes) / stage_bytes >= 4 ? 2 : 1;
    if (inorm) p.nstg = 1;     // the tile is normalised in place in its staging buffer
    const uint32_t budget = smem_cap - fixed - halo_bytes - p.nstg * p.staging_bytes;
    p.stages = budget / stage_bytes;
    if (p.stages > kRingStages) p.stages = kRingStages;
    { const char* e = getenv("TRIS_GEMM_STAGES"); if (e && atoi(e) >= 2 && (uint32_t)atoi(e) < p.stages) p.stages = atoi(e); }
    if (p.stages < 2) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: stage too large (%u bytes)", stage_bytes);
    // two epilogue groups on alternating tiles: plain epilogue, two staging buffers, more than one tile per CTA
    p.ng = 1;
    {
        static int ng_env = -1;
        if (ng_env < 0) { const char* e = getenv("TRIS_GEMM_NG"); ng_env = e ? atoi(e) : 2; }
        const long tiles_all = static_cast<long>(p.tiles_tap) * p.tiles_m * p.tiles_n * p.split_k * p.batch;
        const bool stats_ok = !g->stats || g->N <= 1024;
        if (ng_env == 2 && !inorm && g->stats_mode != 1 && !g->mask_sc && p.nstg == 2 && tiles_all > tris::sm_count() && stats_ok) p.ng = 2;
    }
    p.b_stationary = 0;
    if (halo && p.cblocks == 1 && p.tiles_n == 1 && kMaxStages >= 9) {
        // weight-stationary: all nine taps fit next to the halo ring -> loaded once per CTA instead of once per tile
        const uint32_t need = halo_bytes + 9 * stage_bytes;
        uint32_t nst = p.nstg;
        if (fixed + need + nst * p.staging_bytes > smem_cap) nst = 1;
        if (fixed + need + nst * p.staging_bytes <= smem_cap) { p.b_stationary = 1; p.stages = 9; p.nstg = nst; }
    }
    p.ring_bytes = halo_bytes + p.stages * stage_bytes;
    const size_t smem_bytes = fixed + p.ring_bytes + p.nstg * p.staging_bytes;

    p.fd_tiles_n = make_fastdiv(p.tiles_n); p.fd_tiles_m = make_fastdiv(p.tiles_m); p.fd_tiles_tap = make_fastdiv(p.tiles_tap);
    p.fd_split = make_fastdiv(p.split_k); p.fd_tiles_w = make_fastdiv(p.tiles_w); p.fd_tiles_h = make_fastdiv(p.tiles_h);
    p.fd_cblocks = make_fastdiv(p.cblocks > 0 ? p.cblocks : 1); p.fd_tw = make_fastdiv(p.tw > 0 ? p.tw : 1); p.fd_bn = make_fastdiv(bn);
    p.idesc = ptx::umma_idesc(1u, a_mn ? 1u : 0u, b_mn ? 1u : 0u, kBlockM, bn);
    if (p.dbg & 8) {
        static long long* buf = nullptr;
        if (!buf) { const char* e = getenv("TRIS_GEMM_DBGBUF"); buf = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
        p.dbg_buf = buf;
        if (!buf) p.dbg &= ~8;
    }
    p.d = g->d; p.bias = g->bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(g->residual); p.stats = g->stats;
    p.stats_parts = g->stats_parts; p.stats_mode = g->stats_mode;
    p.stats_y = reinterpret_cast<const __nv_bfloat16*>(g->stats_y); p.stats_mu = g->stats_mu;
    p.mask_sc = g->mask_sc; p.mask_sh = g->mask_sh;
    p.tstamp = reinterpret_cast<unsigned long long*>(g->tstamp);
    p.in_gamma = g->in_gamma; p.in_beta = g->in_beta; p.in_mean = g->in_mean; p.in_invstd = g->in_invstd;
    p.in_add = reinterpret_cast<const __nv_bfloat16*>(g->in_add);
    p.in_eps = g->in_eps; p.in_mix = g->in_mix == 0.f ? 1.f : g->in_mix; p.in_relu = g->in_relu;
    p.res_bits = reinterpret_cast<const unsigned char*>(g->res_bits);
    if (g->res_bits && (!g->residual || g->N % 32 || batch > 1 || conv)) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: res_bits needs a residual, N %% 32 == 0, plain 2-D mode");
    p.d_pre = reinterpret_cast<__nv_bfloat16*>(g->d_pre);
    p.dact_src = reinterpret_cast<const __nv_bfloat16*>(g->dact_src);
    p.ldd = g->ldd; p.act = g->act; p.out_f32 = g->out_dtype == TRIS_DT_F32; p.atomic = g->atomic;
    p.res_f32 = g->residual_f32;
    if (g->residual_f32 && (!g->residual || g->res_bits || g->N % 4 || g->a_mode != TRIS_OP_K2D || g->b_mode != TRIS_OP_K2D || batch > 1 ||
                            g->stats || g->dact_src))
        return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: residual_f32 needs a residual, plain K-major 2-D operands, no res_bits / stats / dact, N %% 4 == 0");
    p.scale = g->scale == 0.f ? 1.f : g->scale;

    // ---- tensor maps
    const CUtensorMap *ma = nullptr, *mb = nullptr;
    if (conv) {
        const int C = a_ch;
        if (C % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: conv channels %d %% 8", C);
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.img_w, (uint64_t)p.img_h, (uint64_t)p.img_n};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.img_w * C * 2, (uint64_t)p.img_h * p.img_w * C * 2};
        uint32_t box[4] = {64, (uint32_t)(p.tw + (halo ? 2 : 0)), (uint32_t)(p.th + (halo ? 2 : 0)), 1};
        ma = tris::tensor_map_bf16(g->a, 4, dims, str, box);
    } else if (g->a_mode == TRIS_OP_K2D) {
        if (g->lda % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: lda %d %% 8", g->lda);
        uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->M, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)g->lda * 2, (uint64_t)g->a_batch_stride * 2};
        uint32_t box[3] = {64, kBlockM, 1};
        ma = tris::tensor_map_bf16(g->a, p.a_batched ? 3 : 2, dims, str, box);
    } else {
        if (g->lda % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: lda %d %% 8", g->lda);
        uint64_t dims[3] = {(uint64_t)g->M, (uint64_t)g->K, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)g->lda * 2, (uint64_t)g->a_batch_stride * 2};
        uint32_t box[3] = {64, (uint32_t)p.bk, 1};
        ma = tris::tensor_map_bf16(g->a, p.a_batched ? 3 : 2, dims, str, box);
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
        uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->N, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)g->ldb * 2, (uint64_t)g->b_batch_stride * 2};
        uint32_t box[3] = {64, (uint32_t)bn, 1};
        mb = tris::tensor_map_bf16(g->b, p.b_batched ? 3 : 2, dims, str, box);
    } else {
        if (g->ldb % 8) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: ldb %d %% 8", g->ldb);
        // [K rows, ldb] row-major; for a conv dgrad the contiguous dim holds taps x N.
        const uint64_t inner = (conv && !wgrad) ? (uint64_t)p.taps * g->b_tap_stride : (uint64_t)g->N;
        const uint64_t outer = (conv && !wgrad) ? (uint64_t)a_ch : (uint64_t)g->K;
        uint64_t dims[3] = {inner, outer, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)g->ldb * 2, (uint64_t)g->b_batch_stride * 2};
        uint32_t box[3] = {64, (uint32_t)p.bk, 1};
        mb = tris::tensor_map_bf16(g->b, p.b_batched ? 3 : 2, dims, str, box);
    }
    if (!mb) return TRIS_ERR_SHAPE;
    // ---- output tensor map (TMA store / reduce-add, 128-byte swizzled staging groups)
    const CUtensorMap* md = nullptr;
    {
        const int esz = out_f32 ? 4 : 2;
        if ((g->ldd * esz) % 16) return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: ldd %d not 16-byte aligned", g->ldd);
        if (conv && !wgrad) {
            uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)p.img_w, (uint64_t)p.img_h, (uint64_t)p.img_n};
            uint64_t str[3] = {(uint64_t)g->ldd * 2, (uint64_t)p.img_w * g->ldd * 2, (uint64_t)p.img_h * p.img_w * g->ldd * 2};
            uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
            md = tris::tensor_map_bf16(g->d, 4, dims, str, box, 2);
        } else if (wgrad && p.split_ws) {
            uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)p.taps, (uint64_t)g->M, (uint64_t)p.split_k};
            uint64_t str[3] = {(uint64_t)g->N * 4, (uint64_t)ws_w * 4, (uint64_t)g->M * ws_w * 4};
            uint32_t box[4] = {32, 1, (uint32_t)kBlockM, 1};
            md = tris::tensor_map_bf16(g->splitk_ws, 4, dims, str, box, 4);
        } else if (wgrad) {
            uint64_t dims[3] = {(uint64_t)g->N, (uint64_t)p.taps, (uint64_t)g->M};
            uint64_t str[2] = {(uint64_t)g->N * 4, (uint64_t)g->ldd * 4};
            uint32_t box[3] = {32, 1, (uint32_t)kBlockM};
            md = tris::tensor_map_bf16(g->d, 3, dims, str, box, 4);
        } else if (p.split_ws) {
            uint64_t dims[3] = {(uint64_t)g->N, (uint64_t)g->M, (uint64_t)p.split_k};
            uint64_t str[2] = {(uint64_t)ws_w * 4, (uint64_t)g->M * ws_w * 4};
            uint32_t box[3] = {32, (uint32_t)kBlockM, 1};
            md = tris::tensor_map_bf16(g->splitk_ws, 3, dims, str, box, 4);
        } else {
            uint64_t dims[3] = {(uint64_t)g->N, (uint64_t)g->M, (uint64_t)batch};
            uint64_t str[2] = {(uint64_t)g->ldd * esz, (uint64_t)g->d_batch_stride * esz};
            uint32_t box[3] = {(uint32_t)(128 / esz), (uint32_t)kBlockM, 1};
            if (batch > 1 && ((g->d_batch_stride * esz) % 16 || g->d_batch_stride == 0))
                return tris::fail(TRIS_ERR_ALIGN, "tris_gemm: d_batch_stride must be non-zero and 16-byte aligned");
            md = tris::tensor_map_bf16(g->d, batch > 1 ? 3 : 2, dims, str, box, esz);
        }
    }
    if (!md) return TRIS_ERR_SHAPE;
    const CUtensorMap* me = md;
    if (inorm) {
        uint64_t dims[3] = {(uint64_t)g->N, (uint64_t)g->M, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)g->ldd * 2, (uint64_t)g->d_batch_stride * 2};
        uint32_t box[3] = {64, (uint32_t)kBlockM, 1};
        me = tris::tensor_map_bf16(g->d_norm, 3, dims, str, box, 2);
        if (!me) return TRIS_ERR_SHAPE;
    }

    int am = LD_K2D, bm = LD_K2D;
    if (halo) am = LD_CONV_HALO; else if (conv && !wgrad) am = LD_CONV; else if (wgrad) am = LD_CONV_WG; else if (g->a_mode == TRIS_OP_MN2D) am = LD_MN2D;
    if (wgrad) bm = LD_CONV_WG; else if (g->b_mode == TRIS_OP_MN2D) bm = (conv ? LD_MN2D_TAPS : LD_MN2D);
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const KParams);
    KernelFn fn = nullptr;
    const int epi = inorm ? 2 : ((g->stats_mode == 1 || g->mask_sc != nullptr) ? 1 : 0);
#define TRIS_PICK(A_, B_) (epi == 1 ? static_cast<KernelFn>(tris_umma_gemm_kernel<A_, B_, 1>) : static_cast<KernelFn>(tris_umma_gemm_kernel<A_, B_, 0>))
    if (am == LD_K2D && bm == LD_K2D && g->residual_f32) fn = tris_umma_gemm_kernel<LD_K2D, LD_K2D, 3>;
    else if (am == LD_K2D && bm == LD_K2D) fn = epi == 2 ? static_cast<KernelFn>(tris_umma_gemm_kernel<LD_K2D, LD_K2D, 2>) : TRIS_PICK(LD_K2D, LD_K2D);
    else if (am == LD_K2D && bm == LD_MN2D) fn = TRIS_PICK(LD_K2D, LD_MN2D);
    else if (am == LD_MN2D && bm == LD_MN2D) fn = tris_umma_gemm_kernel<LD_MN2D, LD_MN2D, 0>;
    else if (am == LD_MN2D && bm == LD_K2D) fn = tris_umma_gemm_kernel<LD_MN2D, LD_K2D, 0>;
    else if (am == LD_CONV && bm == LD_K2D) fn = TRIS_PICK(LD_CONV, LD_K2D);
    else if (am == LD_CONV && bm == LD_MN2D_TAPS) fn = TRIS_PICK(LD_CONV, LD_MN2D_TAPS);
    else if (am == LD_CONV_WG && bm == LD_CONV_WG) fn = tris_umma_gemm_kernel<LD_CONV_WG, LD_CONV_WG, 0>;
    else if (am == LD_CONV_HALO && bm == LD_K2D && !epi) fn = tris_umma_gemm_kernel<LD_CONV_HALO, LD_K2D, 0>;
    else if (am == LD_CONV_HALO && bm == LD_MN2D_TAPS && !epi) fn = tris_umma_gemm_kernel<LD_CONV_HALO, LD_MN2D_TAPS, 0>;
    else return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: unsupported operand mode combination a=%d b=%d", g->a_mode, g->b_mode);
#undef TRIS_PICK
    if (epi && (am == LD_MN2D || am == LD_CONV_WG)) return tris::fail(TRIS_ERR_SHAPE, "tris_gemm: BatchNorm-backward epilogue needs a K-major / conv A operand");
    static bool attr_set[8][8][4] = {};
    const int inst = g->residual_f32 ? 3 : epi;
    if (!attr_set[am][bm][inst]) {
        TRIS_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[am][bm][inst] = true;
    }
    const int total_tiles = p.tiles_tap * p.tiles_m * p.tiles_n * p.split_k * p.batch;
    int ctas = tris::sm_count();
    if (g->max_ctas > 0 && g->max_ctas < ctas) ctas = g->max_ctas;
    if (total_tiles < ctas) ctas = total_tiles;
    if (g->stats && ctas > g->stats_parts) ctas = g->stats_parts;   // one partial-statistics row per CTA
    fn<<<ctas, kThreads, smem_bytes, stream>>>(*ma, *mb, *md, *me, p);
    TRIS_LAUNCH_OK("tris_umma_gemm_kernel");
    if (p.split_ws && !g->defer_reduce) {
        const long total4 = static_cast<long>(g->M) * ws_w / 4;
        long blocks = (total4 + 255) / 256;
        if (blocks > 2L * tris::sm_count()) blocks = 2L * tris::sm_count();
        splitk_reduce_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(g->splitk_ws, reinterpret_cast<float*>(g->d), g->M, ws_w, g->ldd,
                                                                          p.split_k, g->atomic);
        TRIS_LAUNCH_OK("splitk_reduce_kernel");
    }
    return TRIS_OK;
}
