// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved through the runtime, so the library does
// not link libcuda) and a small per-process cache keyed on the full map description.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace vitk {

struct TmapKey {
    const void* ptr;
    uint64_t dims[5];
    uint64_t strides[4];  // bytes, dims 1..rank-1
    uint32_t box[5];
    uint32_t rank;
    uint32_t swizzle;  // CUtensorMapSwizzle
    uint32_t dtype;    // CUtensorMapDataType
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ p[i]) * 1099511628211ull;
        return static_cast<size_t>(h);
    }
};

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// Returns 0 on success. dims/box innermost first; strides[i] = byte stride of dim i+1.
inline int make_tmap(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, CUtensorMapSwizzle swizzle, CUtensorMapDataType dtype) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr;
    key.rank = rank;
    key.swizzle = static_cast<uint32_t>(swizzle);
    key.dtype = static_cast<uint32_t>(dtype);
    for (uint32_t i = 0; i < rank; ++i) {
        key.dims[i] = dims[i];
        key.box[i] = box[i];
        if (i + 1 < rank) key.strides[i] = strides[i];
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    auto fn = tmap_encode_fn();
    if (fn == nullptr) return -10;
    cuuint64_t gdims[5], gstrides[4];
    cuuint32_t gbox[5], estr[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstrides[i] = strides[i];
    }
    CUresult r = fn(out, dtype, rank, const_cast<void*>(ptr), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -11;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 65536) cache.clear();
        cache.emplace(key, *out);
    }
    return 0;
}

// 2-D bf16 map over a row-major [outer, inner] view with a row pitch in elements.
inline int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                             uint32_t box_inner, uint32_t box_outer) {
    uint64_t dims[2] = {inner, outer};
    uint64_t strides[1] = {pitch_elems * 2};
    uint32_t box[2] = {box_inner, box_outer};
    return make_tmap(out, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace vitk
