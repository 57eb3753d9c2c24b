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

// 2-D map over a row-major [rows, cols] OUTPUT (bf16 or fp32) for TMA stores of [box_rows x 64 bytes] SWIZZLE_64B boxes.
inline int make_tmap_2d_store(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t cols, uint64_t rows,
                              uint64_t pitch_elems, uint32_t box_rows) {
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {pitch_elems * (uint64_t)elem_bytes};
    uint32_t box[2] = {(uint32_t)(64 / elem_bytes), box_rows};
    return make_tmap(out, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B,
                     elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

// 3-D bf16 map over a token-major [B, N, cols] OUTPUT for TMA stores of one image's [box_rows x 64 columns] SWIZZLE_128B
// tile: rows past the end of the image (>= N) are clipped by the map instead of landing in the next image.
inline int make_tmap_3d_tok_store(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t N, uint64_t B,
                                  uint32_t box_rows) {
    uint64_t dims[3] = {cols, N, B};
    uint64_t strides[2] = {cols * 2, cols * N * 2};
    uint32_t box[3] = {64, box_rows, 1};
    return make_tmap(out, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
// 4-D bf16 map over a batched row-major operand: logical dims (inner | rows, batch_h, batch_b) with element strides
// (1 | pitch, stride_h, stride_b). The three outer dims are emitted in order of increasing stride (the driver wants
// each stride to be a multiple of the previous one); perm[i] tells the kernel which logical coordinate
// (0 = row, 1 = batch_h, 2 = batch_b) goes into map dimension 1+i.
inline int make_tmap_4d_bf16(CUtensorMap* out, int perm[3], const void* ptr, uint64_t inner, uint64_t rows, uint64_t nh,
                             uint64_t nb, uint64_t pitch, uint64_t stride_h, uint64_t stride_b, uint32_t box_inner,
                             uint32_t box_rows) {
    uint64_t ext[3] = {rows, nh, nb};
    uint64_t str[3] = {pitch, stride_h, stride_b};
    int order[3] = {0, 1, 2};
    // extent-1 dims carry no addressing: give them a harmless stride and sort them last
    uint64_t maxs = pitch * rows;
    for (int i = 1; i < 3; ++i)
        if (ext[i] > 1 && str[i] * ext[i] > maxs) maxs = str[i] * ext[i];
    for (int i = 1; i < 3; ++i)
        if (ext[i] <= 1) str[i] = maxs;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (str[order[j]] < str[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    uint64_t dims[4] = {inner, ext[order[0]], ext[order[1]], ext[order[2]]};
    uint64_t strides[3] = {str[order[0]] * 2, str[order[1]] * 2, str[order[2]] * 2};
    uint32_t box[4] = {box_inner, order[0] == 0 ? box_rows : 1u, order[1] == 0 ? box_rows : 1u,
                       order[2] == 0 ? box_rows : 1u};
    for (int i = 0; i < 3; ++i) perm[i] = order[i];
    return make_tmap(out, ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

// SMs the persistent kernels size their grids for. vitk_set_sm_limit() lowers it (data-parallel runs leave a few SMs
// to the NCCL kernels: a statically scheduled persistent CTA that has to wait for an SM doubles its kernel's time).
inline int& sm_limit_ref() {
    static int limit = 0;  // 0 = no limit
    return limit;
}
inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    const int lim = sm_limit_ref();
    return (lim > 0 && lim < n) ? lim : n;
}

}  // namespace vitk
