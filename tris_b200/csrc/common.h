// Host-side helpers shared by the C-ABI translation units: error channel + TMA tensor-map factory.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tris_sm100.h"

namespace tris {

int fail(int code, const char* fmt, ...);
int sm_count();

#define TRIS_CUDA_OK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return ::tris::fail(TRIS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define TRIS_LAUNCH_OK(name)                                                                     \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) return ::tris::fail(TRIS_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
    } while (0)

// bf16 tensor map, SWIZZLE_128B, inner box = 64 elements (128 bytes).  dims/strides innermost first;
// strides[i] is the byte stride of dimension i+1.  Cached per (ptr, geometry); thread-safe.
const CUtensorMap* tensor_map_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, int elem_bytes = 2);

}  // namespace tris
