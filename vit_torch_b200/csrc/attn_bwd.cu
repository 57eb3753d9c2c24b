// Fused attention backward on tcgen05 (flash-style recompute; nothing of size [B,H,N,N] is ever stored).
//
//   P = exp2(S*scale*log2e - lse2),  dP = dO V^T,  dS = scale * P o (dP - delta),  delta_i = sum_e dO[i,e] O[i,e]
//   dQ = dS K        dK = dS^T Q        dV = P^T dO
// (the reference gets these from autograd over models/swin.py:119-144-style eager attention.)
//
// Two kernels, both deterministic (no atomics, no fp32 workspace):
//   attn_bwd_dq_kernel : CTA = (128 query rows, head, batch), loops over 64-row K/V tiles; dQ accumulates in TMEM.
//                        Also produces delta (written to global for the second kernel).
//   attn_bwd_dkv_kernel: CTA = (128 key rows,   head, batch), loops over 64-row Q/dO tiles; dK, dV accumulate in TMEM.
// S and dP are recomputed in both (7 tile-GEMMs instead of 5) in exchange for no dQ atomics / conversion pass.
// Inputs are read in place from the qkv Linear output [B,N,3,H,d] and dO [B,N,H,d] via rank-2 TMA maps (rows of a tile
// past the image end belong to the next image and are masked; see make_tok_tmap2d); gradients are
// written in place into dqkv [B,N,3,H,d] (the layout the qkv dgrad/wgrad GEMMs consume).
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int AB_THREADS = 192;       // warps 0-3 elementwise (thread == TMEM lane), warp 4 TMA, warp 5 MMA
constexpr int AB_T128 = 128 * 128;    // [128 rows x 64 bf16] swizzled tile bytes
constexpr int AB_T64 = 64 * 128;      // [64 rows x 64 bf16] swizzled tile bytes
constexpr uint32_t AB_TMEM_COLS = 256;

struct AttnBwdArgs {
    int B, H, N, D;
    float scale, scale_log2;
    const __nv_bfloat16* out;   // forward output O [B*N, D]
    const __nv_bfloat16* dout;  // dO [B*N, D]
    const float* lse2;          // [B,H,N]
    float* delta;               // [B,H,N]
    __nv_bfloat16* dqkv;        // [B*N, 3D]
    float* dbias;               // optional fp32 [3D], += column sums of dqkv (gradient of the qkv Linear bias)
    long long* trace;           // instrumented build only (VITK_TRACE), else nullptr
    int dbg_skip;               // instrumented build only: 1 = skip the dQ kernel, 2 = skip the dK/dV kernel
    int pf_dist;                // two-group kernels: CTA i pulls the tiles of CTA i + pf_dist into L2 (0 = off)
};

#ifdef VITK_TRACE
extern long long* g_attn_trace;  // attn_fwd.cu
#endif

// += column sums of one [rows of this warp x 16 columns] chunk of a gradient tile (invalid rows contribute zero)
__device__ __forceinline__ void ab_colsum_chunk(float* dst16, const uint32_t* r, bool row_ok, int lane) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = row_ok ? __uint_as_float(r[i]) : 0.f;
    const float tot = warp_colsum16(v, lane);
    if ((lane & 1) == 0) atomicAdd(dst16 + warp_colsum16_col(lane), tot);
}

// write 32 consecutive bf16 columns (chunk c of a 64-column K-major swizzled tile row)
__device__ __forceinline__ void store_row_chunk_sw128(uint8_t* tile_row, int sw, int c, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint4 pk = make_uint4(pack_bf16(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16(v[u * 8 + 2], v[u * 8 + 3]),
                                    pack_bf16(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16(v[u * 8 + 6], v[u * 8 + 7]));
        const int unit = c * 4 + u;
        *reinterpret_cast<uint4*>(tile_row + ((unit ^ sw) << 4)) = pk;
    }
}

// write 8 consecutive bf16 columns (16-byte unit `unit` of a 64-column K-major swizzled tile row)
__device__ __forceinline__ void store_row_unit_sw128(uint8_t* tile_row, int sw, int unit, const float* v) {
    *reinterpret_cast<uint4*>(tile_row + ((unit ^ sw) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// ------------------------------------------------------------------------------------------------------------------
// dQ kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int DQ_SMEM_Q = 0;                          // [128 x 64]
constexpr int DQ_SMEM_DO = DQ_SMEM_Q + AB_T128;       // [128 x 64]
constexpr int DQ_SMEM_K = DQ_SMEM_DO + AB_T128;       // 2 stages [64 x 64]
constexpr int DQ_SMEM_V = DQ_SMEM_K + 2 * AB_T64;     // 2 stages [64 x 64]
constexpr int DQ_SMEM_DS = DQ_SMEM_V + 2 * AB_T64;    // [128 x 64] dS (K-major A operand)
constexpr int DQ_SMEM_BAR = DQ_SMEM_DS + AB_T128;
constexpr int DQ_SMEM_BYTES = DQ_SMEM_BAR + 256;

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_BAR);
    uint64_t* qdo_full = bars + 0;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* sdp_full = bars + 5;
    uint64_t* sdp_free = bars + 6;
    uint64_t* ds_full = bars + 7;
    uint64_t* ds_free = bars + 8;
    uint64_t* dq_full = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * 128;
    const int nkv = (a.N + 63) / 64;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO128);
        mbar_init(qdo_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 128);
        mbar_init(ds_full, 128);
        mbar_init(ds_free, 1);
        mbar_init(dq_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dq = tmem_base + 128;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(qdo_full, 2 * AB_T128);
            tma_load_2d(smem + DQ_SMEM_Q, &tmQKV128, qdo_full, h * HD, b * a.N + q0);
            tma_load_2d(smem + DQ_SMEM_DO, &tmDO128, qdo_full, h * HD, b * a.N + q0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * AB_T64);
                tma_load_2d(smem + DQ_SMEM_K + s * AB_T64, &tmQKV64, &kv_full[s], (a.H + h) * HD, b * a.N + j * 64);
                tma_load_2d(smem + DQ_SMEM_V + s * AB_T64, &tmQKV64, &kv_full[s], (2 * a.H + h) * HD, b * a.N + j * 64);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + DQ_SMEM_Q), do_addr = smem_u32(smem + DQ_SMEM_DO);
            const uint32_t ds_addr = smem_u32(smem + DQ_SMEM_DS);
            auto issue_sdp = [&](int j) {
                const int s = j & 1;
                const int valid = min(64, a.N - j * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait(&kv_full[s], (j >> 1) & 1);
                if (j > 0) mbar_wait(sdp_free, (j - 1) & 1);
                tc_fence_after_sync();
                const uint64_t qd = make_smem_desc_sw128(q_addr, 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(do_addr, 0, 1024);
                const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + DQ_SMEM_K + s * AB_T64), 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + DQ_SMEM_V + s * AB_T64), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_s, qd + 2 * k, kd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dp, dod + 2 * k, vd + 2 * k, idesc, k > 0);
                umma_commit(sdp_full);
            };
            mbar_wait(qdo_full, 0);
            issue_sdp(0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                if (j + 1 < nkv) issue_sdp(j + 1);
                const int valid = min(64, a.N - j * 64);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait(ds_full, j & 1);
                tc_fence_after_sync();
                // dQ[128, HD] += dS[128, kv] * K_j[kv, HD]: A = dS (K-major), B = K tile read MN-major
                constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t k_addr = smem_u32(smem + DQ_SMEM_K + s * AB_T64);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t ad = make_smem_desc_sw128(ds_addr + k * 32, 0, 1024);
                    const uint64_t bd = make_smem_desc_sw128(k_addr + k * 2048, 64 * 128, 1024);
                    umma_bf16(tmem_dq, ad, bd, idesc_dq, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(ds_free);
                umma_commit(&kv_empty[s]);
            }
            umma_commit(dq_full);
        }
    } else {
        const int row = warp * 32 + lane;
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        const int n = q0 + row;
        const bool row_ok = n < a.N;
        // delta = rowsum(dO o O) for this (b, h, n); also published for the dK/dV kernel
        float delta = 0.f, lse2 = 0.f;
        if (row_ok) {
            const __nv_bfloat16* op = a.out + ((long long)b * a.N + n) * a.D + h * HD;
            const __nv_bfloat16* dop = a.dout + ((long long)b * a.N + n) * a.D + h * HD;
#pragma unroll
            for (int u = 0; u < HD / 8; ++u) {
                const uint4 x = *reinterpret_cast<const uint4*>(op + u * 8);
                const uint4 y = *reinterpret_cast<const uint4*>(dop + u * 8);
                delta += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x);
                delta += bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y);
                delta += bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z);
                delta += bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
            }
            const long long si = ((long long)b * a.H + h) * a.N + n;
            lse2 = a.lse2[si];
            a.delta[si] = delta;
        }
        uint8_t* ds_row = smem + DQ_SMEM_DS + row * 128;
        const int sw = row & 7;
        const float rscale = row_ok ? a.scale : 0.f;  // rows beyond N contribute nothing
        for (int j = 0; j < nkv; ++j) {
            const int valid = min(64, a.N - j * 64);
            mbar_wait(sdp_full, j & 1);
            tc_fence_after_sync();
            if (j > 0) mbar_wait(ds_free, (j - 1) & 1);  // previous dQ MMA finished reading the dS tile
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t sr[32], dpr[32];
                tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, sr);
                tmem_ld_32x32b_x32(tmem_dp + lane_off + c * 32, dpr);
                tmem_ld_wait();
                float ds[32];
                if (c * 32 + 32 <= valid) {  // warp-uniform: no per-element masking in fully valid chunks
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = ex2_approx(fmaf(__uint_as_float(sr[i]), a.scale_log2, -lse2));
                        ds[i] = p * (__uint_as_float(dpr[i]) - delta) * rscale;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = ex2_approx(fmaf(__uint_as_float(sr[i]), a.scale_log2, -lse2));
                        const float v = p * (__uint_as_float(dpr[i]) - delta) * rscale;
                        ds[i] = ((c * 32 + i) < valid) ? v : 0.f;
                    }
                }
                store_row_chunk_sw128(ds_row, sw, c, ds);
            }
            tc_fence_before_sync();
            mbar_arrive(sdp_free);
            fence_proxy_async_smem();
            mbar_arrive(ds_full);
        }
        mbar_wait(dq_full, 0);
        tc_fence_after_sync();
        // tcgen05.ld is warp-collective (.sync.aligned): issue it unconditionally, predicate only the stores
        __nv_bfloat16* dst = a.dqkv + ((long long)b * a.N + n) * (3LL * a.D) + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_dq + lane_off + c * 16, r);
            tmem_ld_wait();
            if (a.dbias != nullptr) ab_colsum_chunk(a.dbias + h * HD + c * 16, r, row_ok, lane);
            if (row_ok) {
                const float* f = reinterpret_cast<const float*>(r);
                st_v4(dst + c * 16, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                               pack_bf16(f[6], f[7])));
                st_v4(dst + c * 16 + 8, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]),
                                                   pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc<AB_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// dK / dV kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int DKV_SMEM_K = 0;                          // [128 x 64]
constexpr int DKV_SMEM_V = DKV_SMEM_K + AB_T128;       // [128 x 64]
constexpr int DKV_SMEM_Q = DKV_SMEM_V + AB_T128;       // 2 stages [64 x 64]
constexpr int DKV_SMEM_DO = DKV_SMEM_Q + 2 * AB_T64;   // 2 stages [64 x 64]
constexpr int DKV_SMEM_PT = DKV_SMEM_DO + 2 * AB_T64;  // P^T  [128 kv x 64 q]
constexpr int DKV_SMEM_DST = DKV_SMEM_PT + AB_T128;    // dS^T [128 kv x 64 q]
constexpr int DKV_SMEM_STAT = DKV_SMEM_DST + AB_T128;  // [2 stages][lse2[64], delta[64]] fp32
constexpr int DKV_SMEM_BAR = DKV_SMEM_STAT + 2 * 128 * 4;
constexpr int DKV_SMEM_BYTES = DKV_SMEM_BAR + 256;

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                    const __grid_constant__ CUtensorMap tmDO64, const AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_SMEM_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* qdo_full = bars + 1;   // [2]
    uint64_t* qdo_empty = bars + 3;  // [2]
    uint64_t* st_full = bars + 5;    // S^T / dP^T ready in TMEM
    uint64_t* st_free = bars + 6;
    uint64_t* pds_full = bars + 7;   // P^T / dS^T written to smem
    uint64_t* pds_free = bars + 8;
    uint64_t* dkv_full = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* stat = reinterpret_cast<float*>(smem + DKV_SMEM_STAT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int kv0 = blockIdx.x * 128;
    const int nq = (a.N + 63) / 64;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO64);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
        }
        mbar_init(st_full, 1);
        mbar_init(st_free, 128);
        mbar_init(pds_full, 128);
        mbar_init(pds_free, 1);
        mbar_init(dkv_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 64, tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 192;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * AB_T128);
            tma_load_2d(smem + DKV_SMEM_K, &tmQKV128, kv_full, (a.H + h) * HD, b * a.N + kv0);
            tma_load_2d(smem + DKV_SMEM_V, &tmQKV128, kv_full, (2 * a.H + h) * HD, b * a.N + kv0);
            for (int i = 0; i < nq; ++i) {
                const int s = i & 1;
                mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[s], 2 * AB_T64);
                tma_load_2d(smem + DKV_SMEM_Q + s * AB_T64, &tmQKV64, &qdo_full[s], h * HD, b * a.N + i * 64);
                tma_load_2d(smem + DKV_SMEM_DO + s * AB_T64, &tmDO64, &qdo_full[s], h * HD, b * a.N + i * 64);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t k_addr = smem_u32(smem + DKV_SMEM_K), v_addr = smem_u32(smem + DKV_SMEM_V);
            const uint32_t pt_addr = smem_u32(smem + DKV_SMEM_PT), dst_addr = smem_u32(smem + DKV_SMEM_DST);
            auto issue_st = [&](int i) {
                const int s = i & 1;
                const int valid = min(64, a.N - i * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait(&qdo_full[s], (i >> 1) & 1);
                if (i > 0) mbar_wait(st_free, (i - 1) & 1);
                tc_fence_after_sync();
                const uint64_t kd = make_smem_desc_sw128(k_addr, 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(v_addr, 0, 1024);
                const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + DKV_SMEM_Q + s * AB_T64), 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + DKV_SMEM_DO + s * AB_T64), 0, 1024);
                // S^T[kv, q] = K Q^T ; dP^T[kv, q] = V dO^T
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_st, kd + 2 * k, qd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dpt, vd + 2 * k, dod + 2 * k, idesc, k > 0);
                umma_commit(st_full);
            };
            mbar_wait(kv_full, 0);
            issue_st(0);
            for (int i = 0; i < nq; ++i) {
                const int s = i & 1;
                if (i + 1 < nq) issue_st(i + 1);
                const int valid = min(64, a.N - i * 64);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait(pds_full, i & 1);
                tc_fence_after_sync();
                // dV[kv, HD] += P^T[kv, q] dO[q, HD] ; dK[kv, HD] += dS^T[kv, q] Q[q, HD]   (B tiles read MN-major)
                constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t q_addr = smem_u32(smem + DKV_SMEM_Q + s * AB_T64);
                const uint32_t do_addr = smem_u32(smem + DKV_SMEM_DO + s * AB_T64);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    umma_bf16(tmem_dv, make_smem_desc_sw128(pt_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(do_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                    umma_bf16(tmem_dk, make_smem_desc_sw128(dst_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(q_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                }
                umma_commit(pds_free);
                umma_commit(&qdo_empty[s]);
            }
            umma_commit(dkv_full);
        }
    } else {
        const int row = warp * 32 + lane;  // key index within the tile
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        const int kv = kv0 + row;
        const bool row_ok = kv < a.N;
        const bool row_ok_warp = (kv0 + warp * 32 + 31) < a.N;  // every key row of this warp is valid
        uint8_t* pt_row = smem + DKV_SMEM_PT + row * 128;
        uint8_t* dst_row = smem + DKV_SMEM_DST + row * 128;
        const int sw = row & 7;
        const long long stat_base = ((long long)b * a.H + h) * a.N;
        for (int i = 0; i < nq; ++i) {
            const int valid = min(64, a.N - i * 64);
            // stage lse2 / delta of the 64 queries of this tile in smem (double-buffered, broadcast reads below)
            float* st = stat + (i & 1) * 128;
            {
                const int qi = i * 64 + (row & 63);
                float v = 0.f;
                if (qi < a.N) v = (row < 64) ? a.lse2[stat_base + qi] : a.delta[stat_base + qi];
                st[row] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(st_full, i & 1);
            tc_fence_after_sync();
            if (i > 0) mbar_wait(pds_free, (i - 1) & 1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t sr[32], dpr[32];
                tmem_ld_32x32b_x32(tmem_st + lane_off + c * 32, sr);
                tmem_ld_32x32b_x32(tmem_dpt + lane_off + c * 32, dpr);
                tmem_ld_wait();
                float p[32], ds[32];
                if (row_ok_warp && c * 32 + 32 <= valid) {  // warp-uniform fast path
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float l2 = st[c * 32 + q], dl = st[64 + c * 32 + q];
                        p[q] = ex2_approx(fmaf(__uint_as_float(sr[q]), a.scale_log2, -l2));
                        ds[q] = a.scale * p[q] * (__uint_as_float(dpr[q]) - dl);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float l2 = st[c * 32 + q], dl = st[64 + c * 32 + q];
                        const float pv = ex2_approx(fmaf(__uint_as_float(sr[q]), a.scale_log2, -l2));
                        const bool ok = row_ok && (c * 32 + q) < valid;
                        p[q] = ok ? pv : 0.f;
                        ds[q] = ok ? a.scale * pv * (__uint_as_float(dpr[q]) - dl) : 0.f;
                    }
                }
                store_row_chunk_sw128(pt_row, sw, c, p);
                store_row_chunk_sw128(dst_row, sw, c, ds);
            }
            tc_fence_before_sync();
            mbar_arrive(st_free);
            fence_proxy_async_smem();
            mbar_arrive(pds_full);
        }
        mbar_wait(dkv_full, 0);
        tc_fence_after_sync();
        __nv_bfloat16* dkp = a.dqkv + ((long long)b * a.N + kv) * (3LL * a.D) + a.D + h * HD;
        __nv_bfloat16* dvp = dkp + a.D;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
            uint32_t rk[16], rv[16];
            tmem_ld_32x32b_x16(tmem_dk + lane_off + c * 16, rk);
            tmem_ld_32x32b_x16(tmem_dv + lane_off + c * 16, rv);
            tmem_ld_wait();
            if (a.dbias != nullptr) {
                ab_colsum_chunk(a.dbias + a.D + h * HD + c * 16, rk, row_ok, lane);
                ab_colsum_chunk(a.dbias + 2 * a.D + h * HD + c * 16, rv, row_ok, lane);
            }
            if (row_ok) {
                const float* f = reinterpret_cast<const float*>(rk);
                st_v4(dkp + c * 16, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                               pack_bf16(f[6], f[7])));
                st_v4(dkp + c * 16 + 8, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]),
                                                   pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
                const float* g = reinterpret_cast<const float*>(rv);
                st_v4(dvp + c * 16, make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]),
                                               pack_bf16(g[6], g[7])));
                st_v4(dvp + c * 16 + 8, make_uint4(pack_bf16(g[8], g[9]), pack_bf16(g[10], g[11]),
                                                   pack_bf16(g[12], g[13]), pack_bf16(g[14], g[15])));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc<AB_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Two-group variants (default). Same tiles, barriers and MMA sequence as the kernels above, but EIGHT elementwise warps
// per CTA: group g = warp / 4 handles columns [32g, 32g + 32) of every 64-column S / dP tile (TMEM lane = row as
// before). The per-tile dependent chain LDTM -> exp2 -> dS -> pack -> STS, which bounds these latency-limited kernels,
// is half as long per thread, the S / dP columns are released to the MMA warp right after the TMEM load (before the
// math), and twice as many warps cover the barrier round trips. Warps whose 32 rows all lie past the end of the
// image only take part in the barriers. The write-out is split between the groups.
// ------------------------------------------------------------------------------------------------------------------
constexpr int AB2_THREADS = 320;      // warps 0-7 elementwise, warp 8 TMA, warp 9 MMA
constexpr int DQ2_NS = 3;                             // K/V ring stages
constexpr int DQ2_SMEM_Q = 0;                         // [128 x 64]
constexpr int DQ2_SMEM_DO = DQ2_SMEM_Q + AB_T128;     // [128 x 64]
constexpr int DQ2_SMEM_K = DQ2_SMEM_DO + AB_T128;     // NS stages [64 x 64]
constexpr int DQ2_SMEM_V = DQ2_SMEM_K + DQ2_NS * AB_T64;
constexpr int DQ2_SMEM_DS = DQ2_SMEM_V + DQ2_NS * AB_T64;   // [128 x 64] dS (K-major A operand)
constexpr int DQ2_SMEM_BAR = DQ2_SMEM_DS + AB_T128;
constexpr int DQ2_SMEM_BYTES = DQ2_SMEM_BAR + 256;

template <int HD>
__global__ void __launch_bounds__(AB2_THREADS, 2)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                    const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmO128,
                    const __grid_constant__ CUtensorMap tmDQKV, const AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NS = DQ2_NS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ2_SMEM_BAR);
    uint64_t* qdo_full = bars + 0;
    uint64_t* sdp_full = bars + 1;
    uint64_t* sdp_free = bars + 2;
    uint64_t* ds_full = bars + 3;
    uint64_t* ds_free = bars + 4;
    uint64_t* dq_full = bars + 5;
    uint64_t* kv_full = bars + 6;   // [NS]
    uint64_t* kv_empty = kv_full + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * 128;
    const int nkv = (a.N + 63) / 64;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 8 && lane == 0) {
        VITK_TRACE_EV(a.trace, 0);
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO128);
        mbar_init(qdo_full, 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 256);
        mbar_init(ds_full, 256);
        mbar_init(ds_free, 1);
        mbar_init(dq_full, 1);
        fence_mbar_init();
        // first loads go out before the block-wide sync: their latency overlaps the TMEM allocation
        // the forward output tile O goes into the (still unused) dS buffer: delta = rowsum(dO o O) is formed from smem
        mbar_expect_tx(qdo_full, 3 * AB_T128);
        tma_load_2d(smem + DQ2_SMEM_DO, &tmDO128, qdo_full, h * HD, b * a.N + q0);
        tma_load_2d(smem + DQ2_SMEM_DS, &tmO128, qdo_full, h * HD, b * a.N + q0);
        tma_load_2d(smem + DQ2_SMEM_Q, &tmQKV128, qdo_full, h * HD, b * a.N + q0);
        for (int j = 0; j < NS && j < nkv; ++j) {
            mbar_expect_tx(&kv_full[j], 2 * AB_T64);
            tma_load_2d(smem + DQ2_SMEM_K + j * AB_T64, &tmQKV64, &kv_full[j], (a.H + h) * HD, b * a.N + j * 64);
            tma_load_2d(smem + DQ2_SMEM_V + j * AB_T64, &tmQKV64, &kv_full[j], (2 * a.H + h) * HD, b * a.N + j * 64);
        }
    }
    if (warp == 9) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dq = tmem_base + 128;

    if (warp == 8) {
        if (lane == 0) {
            for (int j = NS; j < nkv; ++j) {   // the first NS tiles were issued in the prologue
                const int s = j % NS;
                mbar_wait_backoff(&kv_empty[s], ((j / NS) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * AB_T64);
                tma_load_2d(smem + DQ2_SMEM_K + s * AB_T64, &tmQKV64, &kv_full[s], (a.H + h) * HD, b * a.N + j * 64);
                tma_load_2d(smem + DQ2_SMEM_V + s * AB_T64, &tmQKV64, &kv_full[s], (2 * a.H + h) * HD, b * a.N + j * 64);
            }
        }
        // L2 prefetch for the CTA that will run one wave later on this SM slot (see attn_fwd2_kernel)
        if (lane == 1 && a.pf_dist > 0) {
            const long long lin = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z) +
                                  a.pf_dist;
            if (lin < (long long)gridDim.x * gridDim.y * gridDim.z) {
                const int pq = (int)(lin % gridDim.x), phb = (int)(lin / gridDim.x);
                const int ph = phb % (int)gridDim.y, pb = phb / (int)gridDim.y;
                tma_prefetch_l2_2d(&tmQKV128, ph * HD, pb * a.N + pq * 128);
                tma_prefetch_l2_2d(&tmDO128, ph * HD, pb * a.N + pq * 128);
                tma_prefetch_l2_2d(&tmO128, ph * HD, pb * a.N + pq * 128);
                if (pq == 0) {
                    for (int j = 0; j < nkv; ++j) {
                        tma_prefetch_l2_2d(&tmQKV64, (a.H + ph) * HD, pb * a.N + j * 64);
                        tma_prefetch_l2_2d(&tmQKV64, (2 * a.H + ph) * HD, pb * a.N + j * 64);
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + DQ2_SMEM_Q), do_addr = smem_u32(smem + DQ2_SMEM_DO);
            const uint32_t ds_addr = smem_u32(smem + DQ2_SMEM_DS);
            auto issue_sdp = [&](int j) {
                const int s = j % NS;
                const int valid = min(64, a.N - j * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait_backoff(&kv_full[s], (j / NS) & 1);
                if (j > 0) mbar_wait_backoff(sdp_free, (j - 1) & 1);
                tc_fence_after_sync();
                const uint64_t qd = make_smem_desc_sw128(q_addr, 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(do_addr, 0, 1024);
                const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + DQ2_SMEM_K + s * AB_T64), 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + DQ2_SMEM_V + s * AB_T64), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_s, qd + 2 * k, kd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dp, dod + 2 * k, vd + 2 * k, idesc, k > 0);
                umma_commit(sdp_full);
            };
            VITK_TRACE_EV(a.trace, 1);
            mbar_wait_backoff(qdo_full, 0);
            issue_sdp(0);
            VITK_TRACE_EV(a.trace, 2);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % NS;
                if (j + 1 < nkv) issue_sdp(j + 1);
                const int valid = min(64, a.N - j * 64);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait_backoff(ds_full, j & 1);
                tc_fence_after_sync();
                constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t k_addr = smem_u32(smem + DQ2_SMEM_K + s * AB_T64);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t ad = make_smem_desc_sw128(ds_addr + k * 32, 0, 1024);
                    const uint64_t bd = make_smem_desc_sw128(k_addr + k * 2048, 64 * 128, 1024);
                    umma_bf16(tmem_dq, ad, bd, idesc_dq, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(ds_free);
                umma_commit(&kv_empty[s]);
                if (j < 8) VITK_TRACE_EV(a.trace, 8 + j);
            }
            umma_commit(dq_full);
        }
    } else {
        const int g = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const int n = q0 + row;
        const bool row_ok = n < a.N;
        const bool warp_live = (q0 + wq * 32) < a.N;   // warp-uniform
        uint8_t* ds_row = smem + DQ2_SMEM_DS + row * 128;
        const int sw = row & 7;
        // delta = rowsum(dO o O) of this thread's query row, from the TMA-staged dO and O tiles (no strided global
        // loads); both groups need it, group 0 publishes it for the dK/dV kernel. For d = 48 the tiles' columns 48-63
        // belong to the next head and are left out.
        float delta = 0.f, lse2 = 0.f;
        if (row_ok) lse2 = a.lse2[((long long)b * a.H + h) * a.N + n];
        mbar_wait(qdo_full, 0);
        if (warp_live) {
            const uint8_t* do_row = smem + DQ2_SMEM_DO + row * 128;
#pragma unroll
            for (int u = 0; u < HD / 8; ++u) {
                const uint4 x = *reinterpret_cast<const uint4*>(ds_row + ((u ^ sw) << 4));
                const uint4 y = *reinterpret_cast<const uint4*>(do_row + ((u ^ sw) << 4));
                delta += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x);
                delta += bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y);
                delta += bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z);
                delta += bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
            }
            if (g == 0 && row_ok) a.delta[((long long)b * a.H + h) * a.N + n] = delta;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone has read the O tile: the dS buffer may be written
        if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 3);
        const float rscale = row_ok ? a.scale : 0.f;  // rows beyond N contribute nothing
        for (int j = 0; j < nkv; ++j) {
            const int valid = min(64, a.N - j * 64);
            const int nv = min(32, valid - 32 * g);
            const bool work = warp_live && nv > 0;     // warp-uniform
            mbar_wait(sdp_full, j & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 0 && j < 8) VITK_TRACE_EV(a.trace, 24 + 2 * j);
            uint32_t sr[32], dpr[32];
            if (work) {
                tmem_ld_32x32b_x32(tmem_s + lane_off + g * 32, sr);
                tmem_ld_32x32b_x32(tmem_dp + lane_off + g * 32, dpr);
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            mbar_arrive(sdp_free);                       // S / dP columns go back to the MMA warp before the math
            if (j > 0) mbar_wait(ds_free, (j - 1) & 1);  // previous dQ MMA finished reading the dS tile
            if (work) {
                // 8 columns (one 16-byte unit of the swizzled tile) at a time: keeps the live registers under the
                // 96 that two 320-thread CTAs per SM leave
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float ds[8];
                    if (nv == 32) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float p = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, -lse2));
                            ds[i] = p * (__uint_as_float(dpr[u * 8 + i]) - delta) * rscale;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float p = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, -lse2));
                            const float v = p * (__uint_as_float(dpr[u * 8 + i]) - delta) * rscale;
                            ds[i] = (u * 8 + i < nv) ? v : 0.f;
                        }
                    }
                    store_row_unit_sw128(ds_row, sw, 4 * g + u, ds);
                }
                fence_proxy_async_smem();
            }
            mbar_arrive(ds_full);
            if (threadIdx.x == 0 && j < 8) VITK_TRACE_EV(a.trace, 25 + 2 * j);
        }
        mbar_wait(dq_full, 0);
        tc_fence_after_sync();
        if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 60);
        // d = 64: the dQ tile leaves through a swizzled staging tile (the dS buffer, free once dq_full has fired) and
        // ONE TMA store (rows >= N are clipped by the rank-3 map); d = 48: direct 16-byte stores
        if (warp_live) {
            // tcgen05.ld is warp-collective (.sync.aligned): issue it unconditionally, predicate only the stores
            __nv_bfloat16* dst = a.dqkv + ((long long)b * a.N + n) * (3LL * a.D) + h * HD;
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                if ((c & 1) != g) continue;
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_dq + lane_off + c * 16, r);
                tmem_ld_wait();
                if (a.dbias != nullptr) ab_colsum_chunk(a.dbias + h * HD + c * 16, r, row_ok, lane);
                const float* f = reinterpret_cast<const float*>(r);
                const uint4 lo = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                            pack_bf16(f[6], f[7]));
                const uint4 hi = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                            pack_bf16(f[14], f[15]));
                if constexpr (HD == 64) {
                    *reinterpret_cast<uint4*>(ds_row + (((2 * c) ^ sw) << 4)) = lo;
                    *reinterpret_cast<uint4*>(ds_row + (((2 * c + 1) ^ sw) << 4)) = hi;
                } else if (row_ok) {
                    st_v4(dst + c * 16, lo);
                    st_v4(dst + c * 16 + 8, hi);
                }
            }
            if constexpr (HD == 64) fence_proxy_async_smem();
        }
        if constexpr (HD == 64) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 0) {
                tma_store_3d(&tmDQKV, smem + DQ2_SMEM_DS, h * HD, q0, b);
                tma_store_commit();
                tma_store_wait_read<0>();
            }
        }
    }
    if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 61);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<AB_TMEM_COLS>(tmem_base);
    }
}

constexpr int DKV2_NS = 2;                              // Q / dO ring stages (a third does not fit next to two CTAs per SM)
constexpr int DKV2_SMEM_K = 0;                          // [128 x 64]
constexpr int DKV2_SMEM_V = DKV2_SMEM_K + AB_T128;      // [128 x 64]
constexpr int DKV2_SMEM_Q = DKV2_SMEM_V + AB_T128;      // NS stages [64 x 64]
constexpr int DKV2_SMEM_DO = DKV2_SMEM_Q + DKV2_NS * AB_T64;
constexpr int DKV2_SMEM_PT = DKV2_SMEM_DO + DKV2_NS * AB_T64;   // P^T  [128 kv x 64 q]
constexpr int DKV2_SMEM_DST = DKV2_SMEM_PT + AB_T128;   // dS^T [128 kv x 64 q]
constexpr int DKV2_SMEM_STAT = DKV2_SMEM_DST + AB_T128; // [2 buffers][lse2[64], delta[64]] fp32
constexpr int DKV2_SMEM_BAR = DKV2_SMEM_STAT + 2 * 128 * 4;
constexpr int DKV2_SMEM_BYTES = DKV2_SMEM_BAR + 256;

template <int HD>
__global__ void __launch_bounds__(AB2_THREADS, 2)
attn_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                     const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmDQKV,
                     const AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NS = DKV2_NS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV2_SMEM_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* st_full = bars + 1;    // S^T / dP^T ready in TMEM
    uint64_t* st_free = bars + 2;
    uint64_t* pds_full = bars + 3;   // P^T / dS^T written to smem
    uint64_t* pds_free = bars + 4;
    uint64_t* dkv_full = bars + 5;
    uint64_t* qdo_full = bars + 6;   // [NS]
    uint64_t* qdo_empty = qdo_full + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qdo_empty + NS);
    float* stat = reinterpret_cast<float*>(smem + DKV2_SMEM_STAT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int kv0 = blockIdx.x * 128;
    const int nq = (a.N + 63) / 64;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 8 && lane == 0) {
        VITK_TRACE_EV(a.trace, 0);
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO64);
        mbar_init(kv_full, 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
        }
        mbar_init(st_full, 1);
        mbar_init(st_free, 256);
        mbar_init(pds_full, 256);
        mbar_init(pds_free, 1);
        mbar_init(dkv_full, 1);
        fence_mbar_init();
        mbar_expect_tx(kv_full, 2 * AB_T128);
        tma_load_2d(smem + DKV2_SMEM_K, &tmQKV128, kv_full, (a.H + h) * HD, b * a.N + kv0);
        tma_load_2d(smem + DKV2_SMEM_V, &tmQKV128, kv_full, (2 * a.H + h) * HD, b * a.N + kv0);
        for (int i = 0; i < NS && i < nq; ++i) {
            mbar_expect_tx(&qdo_full[i], 2 * AB_T64);
            tma_load_2d(smem + DKV2_SMEM_Q + i * AB_T64, &tmQKV64, &qdo_full[i], h * HD, b * a.N + i * 64);
            tma_load_2d(smem + DKV2_SMEM_DO + i * AB_T64, &tmDO64, &qdo_full[i], h * HD, b * a.N + i * 64);
        }
    }
    if (warp == 9) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 64, tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 192;

    if (warp == 8) {
        if (lane == 0) {
            for (int i = NS; i < nq; ++i) {   // the first NS tiles were issued in the prologue
                const int s = i % NS;
                mbar_wait_backoff(&qdo_empty[s], ((i / NS) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[s], 2 * AB_T64);
                tma_load_2d(smem + DKV2_SMEM_Q + s * AB_T64, &tmQKV64, &qdo_full[s], h * HD, b * a.N + i * 64);
                tma_load_2d(smem + DKV2_SMEM_DO + s * AB_T64, &tmDO64, &qdo_full[s], h * HD, b * a.N + i * 64);
            }
        }
        // L2 prefetch for the CTA that will run one wave later on this SM slot (see attn_fwd2_kernel)
        if (lane == 1 && a.pf_dist > 0) {
            const long long lin = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z) +
                                  a.pf_dist;
            if (lin < (long long)gridDim.x * gridDim.y * gridDim.z) {
                const int pk = (int)(lin % gridDim.x), phb = (int)(lin / gridDim.x);
                const int ph = phb % (int)gridDim.y, pb = phb / (int)gridDim.y;
                tma_prefetch_l2_2d(&tmQKV128, (a.H + ph) * HD, pb * a.N + pk * 128);
                tma_prefetch_l2_2d(&tmQKV128, (2 * a.H + ph) * HD, pb * a.N + pk * 128);
                if (pk == 0) {
                    for (int i = 0; i < nq; ++i) {
                        tma_prefetch_l2_2d(&tmQKV64, ph * HD, pb * a.N + i * 64);
                        tma_prefetch_l2_2d(&tmDO64, ph * HD, pb * a.N + i * 64);
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            const uint32_t k_addr = smem_u32(smem + DKV2_SMEM_K), v_addr = smem_u32(smem + DKV2_SMEM_V);
            const uint32_t pt_addr = smem_u32(smem + DKV2_SMEM_PT), dst_addr = smem_u32(smem + DKV2_SMEM_DST);
            auto issue_st = [&](int i) {
                const int s = i % NS;
                const int valid = min(64, a.N - i * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait_backoff(&qdo_full[s], (i / NS) & 1);
                if (i > 0) mbar_wait_backoff(st_free, (i - 1) & 1);
                tc_fence_after_sync();
                const uint64_t kd = make_smem_desc_sw128(k_addr, 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(v_addr, 0, 1024);
                const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + DKV2_SMEM_Q + s * AB_T64), 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + DKV2_SMEM_DO + s * AB_T64), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_st, kd + 2 * k, qd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dpt, vd + 2 * k, dod + 2 * k, idesc, k > 0);
                umma_commit(st_full);
            };
            VITK_TRACE_EV(a.trace, 1);
            mbar_wait_backoff(kv_full, 0);
            issue_st(0);
            VITK_TRACE_EV(a.trace, 2);
            for (int i = 0; i < nq; ++i) {
                const int s = i % NS;
                if (i + 1 < nq) issue_st(i + 1);
                const int valid = min(64, a.N - i * 64);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait_backoff(pds_full, i & 1);
                tc_fence_after_sync();
                constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t q_addr = smem_u32(smem + DKV2_SMEM_Q + s * AB_T64);
                const uint32_t do_addr = smem_u32(smem + DKV2_SMEM_DO + s * AB_T64);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    umma_bf16(tmem_dv, make_smem_desc_sw128(pt_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(do_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                    umma_bf16(tmem_dk, make_smem_desc_sw128(dst_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(q_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                }
                umma_commit(pds_free);
                umma_commit(&qdo_empty[s]);
                if (i < 8) VITK_TRACE_EV(a.trace, 8 + i);
            }
            umma_commit(dkv_full);
        }
    } else {
        const int g = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;  // key index within the tile
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const int kv = kv0 + row;
        const bool row_ok = kv < a.N;
        const bool warp_live = (kv0 + wq * 32) < a.N;           // some key row of this warp is valid (warp-uniform)
        const bool row_ok_warp = (kv0 + wq * 32 + 31) < a.N;    // every key row of this warp is valid
        uint8_t* pt_row = smem + DKV2_SMEM_PT + row * 128;
        uint8_t* dst_row = smem + DKV2_SMEM_DST + row * 128;
        const int sw = row & 7;
        // lse2 / delta of the 64 queries of a tile are staged in smem by group 0 (double-buffered, broadcast reads in the
        // math); the values of tile i + 1 are fetched into a register while tile i is processed and stored after it
        const long long stat_base = ((long long)b * a.H + h) * a.N;
        auto load_stat = [&](int i) -> float {
            const int qi = i * 64 + (row & 63);
            float v = 0.f;
            if (g == 0 && qi < a.N) v = (row < 64) ? a.lse2[stat_base + qi] : a.delta[stat_base + qi];
            return v;
        };
        if (g == 0) stat[row] = load_stat(0);
        for (int i = 0; i < nq; ++i) {
            const int valid = min(64, a.N - i * 64);
            const int nv = min(32, valid - 32 * g);
            const bool work = warp_live && nv > 0;     // warp-uniform
            const float* st = stat + (i & 1) * 128;
            float stat_next = 0.f;
            if (i + 1 < nq) stat_next = load_stat(i + 1);
            asm volatile("bar.sync 1, 256;" ::: "memory");   // stat[i & 1] visible; everyone is done with stat[(i+1) & 1]
            const float4* st4 = reinterpret_cast<const float4*>(st + g * 32);        // lse2 of my 32 columns
            const float4* dl4 = reinterpret_cast<const float4*>(st + 64 + g * 32);   // delta
            mbar_wait(st_full, i & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 0 && i < 8) VITK_TRACE_EV(a.trace, 24 + 2 * i);
            uint32_t sr[32], dpr[32];
            if (work) {
                tmem_ld_32x32b_x32(tmem_st + lane_off + g * 32, sr);
                tmem_ld_32x32b_x32(tmem_dpt + lane_off + g * 32, dpr);
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            mbar_arrive(st_free);                          // S^T / dP^T columns go back to the MMA warp before the math
            if (i > 0) mbar_wait(pds_free, (i - 1) & 1);   // previous dV / dK MMA finished reading P^T / dS^T
            if (work) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {      // 8 query columns = one 16-byte unit of the swizzled tiles
                    float p[8], ds[8];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float4 l4 = st4[2 * u + hh], d4 = dl4[2 * u + hh];
                        const float l2[4] = {l4.x, l4.y, l4.z, l4.w};
                        const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
                        if (row_ok_warp && nv == 32) {  // warp-uniform fast path
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c = u * 8 + hh * 4 + q;
                                p[hh * 4 + q] = ex2_approx(fmaf(__uint_as_float(sr[c]), a.scale_log2, -l2[q]));
                                ds[hh * 4 + q] = a.scale * p[hh * 4 + q] * (__uint_as_float(dpr[c]) - dl[q]);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c = u * 8 + hh * 4 + q;
                                const float pv = ex2_approx(fmaf(__uint_as_float(sr[c]), a.scale_log2, -l2[q]));
                                const bool ok = row_ok && c < nv;
                                p[hh * 4 + q] = ok ? pv : 0.f;
                                ds[hh * 4 + q] = ok ? a.scale * pv * (__uint_as_float(dpr[c]) - dl[q]) : 0.f;
                            }
                        }
                    }
                    store_row_unit_sw128(pt_row, sw, 4 * g + u, p);
                    store_row_unit_sw128(dst_row, sw, 4 * g + u, ds);
                }
                fence_proxy_async_smem();
            }
            if (g == 0 && i + 1 < nq) stat[((i + 1) & 1) * 128 + row] = stat_next;   // (its load overlapped the math)
            mbar_arrive(pds_full);
            if (threadIdx.x == 0 && i < 8) VITK_TRACE_EV(a.trace, 25 + 2 * i);
        }
        mbar_wait(dkv_full, 0);
        tc_fence_after_sync();
        if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 60);
        // group 0 writes dK, group 1 writes dV. d = 64: through swizzled staging tiles (the P^T / dS^T buffers, free once
        // dkv_full has fired) and one TMA store each (rows >= N clipped by the rank-3 map); d = 48: direct stores
        uint8_t* stage_row = g ? dst_row : pt_row;
        if (warp_live) {
            const uint32_t tmem_acc = g ? tmem_dv : tmem_dk;
            __nv_bfloat16* dp = a.dqkv + ((long long)b * a.N + kv) * (3LL * a.D) + (1 + g) * a.D + h * HD;
            float* dbias = a.dbias != nullptr ? a.dbias + (1 + g) * a.D + h * HD : nullptr;
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_acc + lane_off + c * 16, r);
                tmem_ld_wait();
                if (dbias != nullptr) ab_colsum_chunk(dbias + c * 16, r, row_ok, lane);
                const float* f = reinterpret_cast<const float*>(r);
                const uint4 lo = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                            pack_bf16(f[6], f[7]));
                const uint4 hi = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                            pack_bf16(f[14], f[15]));
                if constexpr (HD == 64) {
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c) ^ sw) << 4)) = lo;
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c + 1) ^ sw) << 4)) = hi;
                } else if (row_ok) {
                    st_v4(dp + c * 16, lo);
                    st_v4(dp + c * 16 + 8, hi);
                }
            }
            if constexpr (HD == 64) fence_proxy_async_smem();
        }
        if constexpr (HD == 64) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 0) {
                tma_store_3d(&tmDQKV, smem + DKV2_SMEM_PT, a.D + h * HD, kv0, b);
                tma_store_3d(&tmDQKV, smem + DKV2_SMEM_DST, 2 * a.D + h * HD, kv0, b);
                tma_store_commit();
                tma_store_wait_read<0>();
            }
        }
    }
    if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 61);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<AB_TMEM_COLS>(tmem_base);
    }
}

int make_tok_tmap2d(CUtensorMap* out, const void* p, long long rows, long long cols, int box_rows);  // attn_fwd.cu
bool attn_two_groups();  // attn_fwd.cu (VITK_ATTN_WG2)
int attn_prefetch_dist();  // attn_fwd.cu (VITK_ATTN_PREFETCH)

template <int HD>
static int launch_attn_bwd(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& do128,
                           const CUtensorMap& do64, const CUtensorMap& o128, const CUtensorMap& dqkv3,
                           const AttnBwdArgs& a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(attn_bwd_dq2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ2_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(attn_bwd_dkv2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV2_SMEM_BYTES) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr = true;
    }
    dim3 grid((a.N + 127) / 128, a.H, a.B);
    if (attn_two_groups()) {
        if (a.dbg_skip != 1) attn_bwd_dq2_kernel<HD><<<grid, AB2_THREADS, DQ2_SMEM_BYTES, st>>>(q128, q64, do128, o128, dqkv3, a);
        if (cudaGetLastError() != cudaSuccess) return VITK_ERR_CUDA;
        if (a.dbg_skip != 2)
            attn_bwd_dkv2_kernel<HD><<<grid, AB2_THREADS, DKV2_SMEM_BYTES, st>>>(q128, q64, do64, dqkv3, a);
        return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
    }
    attn_bwd_dq_kernel<HD><<<grid, AB_THREADS, DQ_SMEM_BYTES, st>>>(q128, q64, do128, a);
    if (cudaGetLastError() != cudaSuccess) return VITK_ERR_CUDA;
    attn_bwd_dkv_kernel<HD><<<grid, AB_THREADS, DKV_SMEM_BYTES, st>>>(q128, q64, do64, a);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

}  // namespace vitk

using namespace vitk;

static int attn_bwd_impl(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                         float* delta, void* dqkv_bf16, float* dbias, int B, int N, int H, int d, float scale,
                         void* stream) {
    if (B <= 0 || N <= 0 || H <= 0 || !(d == 64 || d == 48)) return VITK_ERR_ARG;
    if (!qkv_bf16 || !out_bf16 || !dout_bf16 || !lse2 || !delta || !dqkv_bf16) return VITK_ERR_ARG;
    CUtensorMap q128, q64, do128, do64, o128, dqkv3;
    const long long rows = (long long)B * N;
    if (make_tmap_3d_tok_store(&dqkv3, dqkv_bf16, 3ull * H * d, (uint64_t)N, (uint64_t)B, 128)) return VITK_ERR_TMAP;
    if (make_tok_tmap2d(&q128, qkv_bf16, rows, 3LL * H * d, 128) || make_tok_tmap2d(&q64, qkv_bf16, rows, 3LL * H * d, 64) ||
        make_tok_tmap2d(&do128, dout_bf16, rows, (long long)H * d, 128) ||
        make_tok_tmap2d(&do64, dout_bf16, rows, (long long)H * d, 64) ||
        make_tok_tmap2d(&o128, out_bf16, rows, (long long)H * d, 128))
        return VITK_ERR_TMAP;
    AttnBwdArgs a;
    a.B = B; a.H = H; a.N = N; a.D = H * d;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out = reinterpret_cast<const __nv_bfloat16*>(out_bf16);
    a.dout = reinterpret_cast<const __nv_bfloat16*>(dout_bf16);
    a.lse2 = lse2;
    a.delta = delta;
    a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
    a.dbias = dbias;
    a.trace = nullptr;
    a.dbg_skip = 0;
    a.pf_dist = attn_prefetch_dist();
#ifdef VITK_TRACE
    a.trace = g_attn_trace;
    if (const char* e = getenv("VITK_ATTN_DBG_SKIP")) a.dbg_skip = atoi(e);
#endif
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d == 64) return launch_attn_bwd<64>(q128, q64, do128, do64, o128, dqkv3, a, st);
    return launch_attn_bwd<48>(q128, q64, do128, do64, o128, dqkv3, a, st);
}

extern "C" int vitk_attn_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                             float* delta, void* dqkv_bf16, int B, int N, int H, int d, float scale, void* stream) {
    return attn_bwd_impl(qkv_bf16, out_bf16, dout_bf16, lse2, delta, dqkv_bf16, nullptr, B, N, H, d, scale, stream);
}

extern "C" int vitk_attn_bwd_ex(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                                float* delta, void* dqkv_bf16, float* dqkv_bias_grad, int B, int N, int H, int d,
                                float scale, void* stream) {
    return attn_bwd_impl(qkv_bf16, out_bf16, dout_bf16, lse2, delta, dqkv_bf16, dqkv_bias_grad, B, N, H, d, scale,
                         stream);
}
