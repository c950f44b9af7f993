// Fused AdamW over the flat parameter buffer (two learning-rate groups, poly-0.9 schedule evaluated on the device so
// the launch can live inside a CUDA graph), also emitting the bf16 shadow copy used by the tensor-core kernels.
// HBM-bound: 16 B/param read + 14 B/param written, 128-bit accesses.
//
// Replaces torch.optim.AdamW + LambdaLR of train_stage1.py:133-144, 368-372 (decoupled weight decay, bias correction,
// eps outside the bias-corrected sqrt, exactly torch's single-tensor update).
#include <cuda_bf16.h>

#include "common.h"

namespace {

struct AdamParams {
    float* p; const float* g; float* m; float* v; __nv_bfloat16* shadow;
    long n, n_group0;
    const int* step;   // completed optimizer steps (device)
    float max_iter, lr0, lr1, beta1, beta2, eps, wd, grad_scale, power;
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
    const int it = *a.step;
    const float sched = powf(fmaxf(1.f - static_cast<float>(it) / a.max_iter, 0.f), a.power);
    const float t = static_cast<float>(it + 1);
    const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
    const float inv_bc2_sqrt = rsqrtf(bc2);
    const long n4 = a.n >> 2;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float lr = ((i << 2) < a.n_group0 ? a.lr0 : a.lr1) * sched;
        float4 p = reinterpret_cast<float4*>(a.p)[i];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g) + i);
        float4 m = reinterpret_cast<float4*>(a.m)[i];
        float4 v = reinterpret_cast<float4*>(a.v)[i];
        float* pp = &p.x; const float* gp = &g4.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = gp[j] * a.grad_scale;
            pp[j] *= 1.f - lr * a.wd;
            mp[j] = a.beta1 * mp[j] + (1.f - a.beta1) * g;
            vp[j] = a.beta2 * vp[j] + (1.f - a.beta2) * g * g;
            const float denom = sqrtf(vp[j]) * inv_bc2_sqrt + a.eps;
            pp[j] -= (lr / bc1) * mp[j] / denom;
        }
        reinterpret_cast<float4*>(a.p)[i] = p;
        reinterpret_cast<float4*>(a.m)[i] = m;
        reinterpret_cast<float4*>(a.v)[i] = v;
        if (a.shadow != nullptr) {
            __nv_bfloat162* s = reinterpret_cast<__nv_bfloat162*>(a.shadow) + 2 * i;
            s[0] = __floats2bfloat162_rn(p.x, p.y);
            s[1] = __floats2bfloat162_rn(p.z, p.w);
        }
    }
}

__global__ void tick_kernel(int* step) { *step += 1; }

// Cross-graph signal for the overlapped gradient all-reduce: a kernel INSIDE the captured step bumps a device counter when
// the gradients behind the image tower are final; a one-thread kernel on the communication stream (outside the graph) holds
// the NCCL all-reduce queued behind it until the counter reaches the step number.  Bounded spin: a lost signal traps.
__global__ void flag_inc_kernel(unsigned* flag) {
    __threadfence();
    atomicAdd(flag, 1u);
}
__global__ void flag_wait_kernel(const unsigned* flag, unsigned target) {
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (static_cast<int>(v - target) >= 0) break;
        if (clock64() - t0 > 20000000000ll) __trap();      // ~10 s
        __nanosleep(200);
    }
}

}  // namespace

extern "C" {

int tris_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, long n, long n_group0, int* step, float max_iter,
                    float lr0, float lr1, float beta1, float beta2, float eps, float wd, float grad_scale, float power,
                    tris_stream_t stream) {
    if (n % 4 || n_group0 % 4) return tris::fail(TRIS_ERR_SHAPE, "tris_adamw_step: n %% 4");
    AdamParams a{p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow), n, n_group0, step, max_iter, lr0, lr1, beta1, beta2, eps, wd,
                 grad_scale, power};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    adamw_kernel<<<tris::sm_count() * 8, 256, 0, s>>>(a);
    TRIS_LAUNCH_OK("adamw_kernel");
    tick_kernel<<<1, 1, 0, s>>>(step);
    TRIS_LAUNCH_OK("tick_kernel");
    return TRIS_OK;
}

int tris_flag_inc(uint32_t* flag, tris_stream_t stream) {
    flag_inc_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(flag);
    TRIS_LAUNCH_OK("flag_inc_kernel");
    return TRIS_OK;
}

int tris_flag_wait(const uint32_t* flag, uint32_t target, tris_stream_t stream) {
    flag_wait_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(flag, target);
    TRIS_LAUNCH_OK("flag_wait_kernel");
    return TRIS_OK;
}

}  // extern "C"
