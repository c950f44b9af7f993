// Error channel, device check and the TMA tensor-map cache of libtris_sm100.so.
#include "common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace tris {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

struct MapKey {
    uint64_t v[16];
    bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t x : k.v) {
            h ^= x;
            h *= 1099511628211ull;
        }
        return static_cast<size_t>(h);
    }
};

const CUtensorMap* tensor_map_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, int elem_bytes) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap*, MapHash> cache;
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.v[0] = reinterpret_cast<uint64_t>(base);
    key.v[1] = static_cast<uint64_t>(rank) | (static_cast<uint64_t>(elem_bytes) << 8);
    for (int i = 0; i < rank; ++i) {
        key.v[2 + i] = dims[i];
        key.v[7 + i] = (i + 1 < rank) ? strides_bytes[i] : 0;
        key.v[11 + i] = box[i];
    }
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    // The key contains the base pointer, so an eager caller whose allocations keep moving would grow the cache without
    // bound: past 32 K entries start over (descriptors are passed to kernels BY VALUE at launch, so none is in use here;
    // the old maps are kept alive in a graveyard because callers may still hold the returned pointer for this launch).
    static std::vector<CUtensorMap*> graveyard;
    if (cache.size() > 32768) {
        for (auto& kv : cache) graveyard.push_back(kv.second);
        cache.clear();
        if (graveyard.size() > 4 * 32768) {
            for (size_t i = 0; i + 32768 < graveyard.size(); ++i) delete graveyard[i];
            graveyard.erase(graveyard.begin(), graveyard.end() - 32768);
        }
    }
    auto fn = encode_fn();
    if (!fn) {
        fail(TRIS_ERR_ARCH, "cuTensorMapEncodeTiled not available from the driver");
        return nullptr;
    }
    CUtensorMap* m = new CUtensorMap;  // passed to kernels by value (__grid_constant__), host memory is enough
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUtensorMapDataType dt = (elem_bytes == 2) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(m, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fail(TRIS_ERR_SHAPE,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu base %p",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
             (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
        delete m;
        return nullptr;
    }
    cache.emplace(key, m);
    return m;
}

}  // namespace tris

extern "C" {

const char* tris_last_error(void) { return tris::g_err; }

int tris_abi_version(void) { return 1; }

int tris_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return tris::fail(TRIS_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) return tris::fail(TRIS_ERR_ARCH, "device is sm_%d%d; libtris_sm100 needs sm_100a (B200)", major, minor);
    if (!tris::encode_fn()) return tris::fail(TRIS_ERR_ARCH, "driver lacks cuTensorMapEncodeTiled");
    return TRIS_OK;
}

}  // extern "C"
