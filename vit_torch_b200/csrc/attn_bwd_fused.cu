// Single-kernel attention backward on tcgen05 (d = 64): S and dP are computed ONCE per (key block, query tile).
//
//   P = exp2(S*scale*log2e - lse2),  dP = dO V^T,  dS = scale * P o (dP - delta),  delta_i = sum_e dO[i,e] O[i,e]
//   dV = P^T dO        dK = dS^T Q        dQ = dS K
// (the reference gets these from autograd over eager attention, models/swin.py:119-144 form.)
//
// The two-kernel backward (attn_bwd.cu) recomputes S / dP and the whole elementwise pass in both kernels: 7 tile-GEMMs
// and 2x the exp / dS work, and ncu shows those kernels latency-bound on exactly that elementwise <-> MMA hand-off
// (tensor pipe 14 %). Here one CTA owns 128 keys of one (batch, head) and walks the 64-row query tiles:
//     S^T = K Q^T, dP^T = V dO^T  (UMMA 128x64x16, double-buffered in TMEM)
//     8 elementwise warps (thread = key row x half of the 64 queries) form P^T and dS^T as bf16 in swizzled smem
//     dV += P^T dO,  dK += dS^T Q  (UMMA 128x64, accumulators resident in TMEM for the whole CTA)
//     dQ_tile = dS K               (UMMA M = 64: A is the SAME dS^T smem tile read MN-major, B = K read MN-major)
// dQ tiles of different key blocks are summed with red.global.add.v4.f32 into an fp32 workspace (4 dedicated warps
// drain the double-buffered dQ accumulator while the next tile is in flight); a conversion kernel writes them as bf16
// into dqkv. delta comes from a small pre-pass. TMEM: 2x64 (S^T) + 2x64 (dP^T) + 64 (dV) + 64 (dK) + 2x64 (dQ) = 512.
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int ABF_E_WARPS = 8;                       // elementwise warps
constexpr int ABF_Q_WARPS = 4;                       // dQ drain warps
constexpr int ABF_W_TMA = ABF_E_WARPS + ABF_Q_WARPS; // warp 12
constexpr int ABF_W_MMA = ABF_W_TMA + 1;             // warp 13: issues S^T / dP^T (and owns the TMEM allocation)
constexpr int ABF_W_ACC = ABF_W_MMA + 1;             // warp 14: issues dV / dK / dQ
constexpr int ABF_THREADS = (ABF_W_ACC + 1) * 32;    // 480
constexpr int ABF_T128 = 128 * 128;                  // [128 rows x 64 bf16] swizzled tile bytes
constexpr int ABF_T64 = 64 * 128;                    // [64 rows x 64 bf16]
constexpr uint32_t ABF_TMEM_COLS = 512;

constexpr int ABF_SMEM_K = 0;
constexpr int ABF_SMEM_V = ABF_SMEM_K + ABF_T128;
constexpr int ABF_QDO_STAGES = 4;                        // Q / dO tile ring: a TMA round trip spans several tiles
constexpr int ABF_SMEM_Q = ABF_SMEM_V + ABF_T128;
constexpr int ABF_SMEM_DO = ABF_SMEM_Q + ABF_QDO_STAGES * ABF_T64;
constexpr int ABF_SMEM_PT = ABF_SMEM_DO + ABF_QDO_STAGES * ABF_T64;   // 2 buffers: P^T  [128 kv x 64 q]
constexpr int ABF_SMEM_DST = ABF_SMEM_PT + 2 * ABF_T128; // 2 buffers: dS^T [128 kv x 64 q]
constexpr int ABF_SMEM_BAR = ABF_SMEM_DST + 2 * ABF_T128;
constexpr int ABF_SMEM_STAT = ABF_SMEM_BAR + 256;        // lse2[Np] | delta[Np] fp32 of the whole (batch, head), Np = 64*nq
constexpr int ABF_MAX_STAT_BYTES = 64 * 1024;
__host__ __device__ constexpr int abf_smem_bytes(int nq) { return ABF_SMEM_STAT + 2 * nq * 64 * 4; }

struct AttnBwdFusedArgs {
    int B, H, N, D;
    float scale, scale_log2;
    const float* lse2;   // [B,H,N]
    const float* delta;  // [B,H,N]
    float* dq_acc;       // fp32 [B*N, D], zero-initialised
    __nv_bfloat16* dqkv; // [B*N, 3D]
};

__device__ __forceinline__ void abf_store_row_chunk(uint8_t* tile_row, int sw, int c, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint4 pk = make_uint4(pack_bf16(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16(v[u * 8 + 2], v[u * 8 + 3]),
                                    pack_bf16(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16(v[u * 8 + 6], v[u * 8 + 7]));
        const int unit = c * 4 + u;
        *reinterpret_cast<uint4*>(tile_row + ((unit ^ sw) << 4)) = pk;
    }
}

__global__ void __launch_bounds__(ABF_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO64, const AttnBwdFusedArgs a) {
    constexpr int HD = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ABF_SMEM_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* qdo_full = bars + 1;    // [ABF_QDO_STAGES]
    uint64_t* qdo_empty = bars + 5;   // [ABF_QDO_STAGES]
    uint64_t* st_full = bars + 9;     // [2] S^T / dP^T of tile i ready in TMEM buffer i&1
    uint64_t* st_free = bars + 11;    // [2]
    uint64_t* pds_full = bars + 13;   // [2] P^T / dS^T written to smem buffer i&1
    uint64_t* pds_free = bars + 15;   // [2]
    uint64_t* dq_full = bars + 17;    // [2]
    uint64_t* dq_free = bars + 19;    // [2]
    uint64_t* dkv_full = bars + 21;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    float* stat = reinterpret_cast<float*>(smem + ABF_SMEM_STAT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int kv0 = blockIdx.x * 128;
    const int nq = (a.N + 63) / 64;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == ABF_W_TMA && lane == 0) {
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO64);
        mbar_init(kv_full, 1);
        for (int i = 0; i < ABF_QDO_STAGES; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&st_full[i], 1);
            mbar_init(&st_free[i], 128);   // one elementwise group (4 warps) per buffer
            mbar_init(&pds_full[i], 128);
            mbar_init(&pds_free[i], 1);
            mbar_init(&dq_full[i], 1);
            mbar_init(&dq_free[i], ABF_Q_WARPS * 32);
        }
        mbar_init(dkv_full, 1);
        fence_mbar_init();
    }
    if (warp == ABF_W_MMA) tmem_alloc<ABF_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 320,
                   tmem_dq = tmem_base + 384;

    if (warp == ABF_W_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * ABF_T128);
            tma_load_2d(smem + ABF_SMEM_K, &tmQKV128, kv_full, (a.H + h) * HD, b * a.N + kv0);
            tma_load_2d(smem + ABF_SMEM_V, &tmQKV128, kv_full, (2 * a.H + h) * HD, b * a.N + kv0);
            for (int i = 0; i < nq; ++i) {
                const int r = i % ABF_QDO_STAGES;
                mbar_wait(&qdo_empty[r], ((i / ABF_QDO_STAGES) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[r], 2 * ABF_T64);
                tma_load_2d(smem + ABF_SMEM_Q + r * ABF_T64, &tmQKV64, &qdo_full[r], h * HD, b * a.N + i * 64);
                tma_load_2d(smem + ABF_SMEM_DO + r * ABF_T64, &tmDO64, &qdo_full[r], h * HD, b * a.N + i * 64);
            }
        }
    } else if (warp == ABF_W_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t k_addr = smem_u32(smem + ABF_SMEM_K), v_addr = smem_u32(smem + ABF_SMEM_V);
            const int kv_valid = min(128, a.N - kv0);
            const int ksteps_kv = (kv_valid + 15) >> 4;
            auto issue_st = [&](int i) {
                const int s = i & 1, r = i % ABF_QDO_STAGES;
                const int valid = min(64, a.N - i * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait(&qdo_full[r], (i / ABF_QDO_STAGES) & 1);
                mbar_wait(&st_free[s], ((i >> 1) & 1) ^ 1);  // tile i-2 has been read out of this TMEM buffer
                tc_fence_after_sync();
                const uint64_t kd = make_smem_desc_sw128(k_addr, 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(v_addr, 0, 1024);
                const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + ABF_SMEM_Q + r * ABF_T64), 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + ABF_SMEM_DO + r * ABF_T64), 0, 1024);
                // S^T[kv, q] = K Q^T ; dP^T[kv, q] = V dO^T
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_st + s * 64, kd + 2 * k, qd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dpt + s * 64, vd + 2 * k, dod + 2 * k, idesc, k > 0);
                umma_commit(&st_full[s]);
            };
            // S^T / dP^T are issued by their own thread, as far ahead as the Q / dO ring and the two TMEM buffers allow:
            // it must not queue behind the wait for the elementwise warps that the dV / dK / dQ issuer sits in
            mbar_wait(kv_full, 0);
            for (int i = 0; i < nq; ++i) issue_st(i);
        }
    } else if (warp == ABF_W_ACC) {
        // ===================== dV / dK / dQ issuer =====================
        if (lane == 0) {
            const uint32_t k_addr = smem_u32(smem + ABF_SMEM_K);
            const int kv_valid = min(128, a.N - kv0);
            const int ksteps_kv = (kv_valid + 15) >> 4;
            for (int i = 0; i < nq; ++i) {
                const int s = i & 1;
                const int valid = min(64, a.N - i * 64);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait(&pds_full[s], (i >> 1) & 1);
                tc_fence_after_sync();
                // dV[kv, HD] += P^T[kv, q] dO[q, HD] ; dK[kv, HD] += dS^T[kv, q] Q[q, HD]   (B tiles read MN-major)
                constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t pt_addr = smem_u32(smem + ABF_SMEM_PT + s * ABF_T128);
                const uint32_t dst_addr = smem_u32(smem + ABF_SMEM_DST + s * ABF_T128);
                const int r = i % ABF_QDO_STAGES;
                const uint32_t q_addr = smem_u32(smem + ABF_SMEM_Q + r * ABF_T64);
                const uint32_t do_addr = smem_u32(smem + ABF_SMEM_DO + r * ABF_T64);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    umma_bf16(tmem_dv, make_smem_desc_sw128(pt_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(do_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                    umma_bf16(tmem_dk, make_smem_desc_sw128(dst_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(q_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                }
                // dQ_tile[q, HD] = dS[q, kv] K[kv, HD]: A = the dS^T tile read MN-major (M = 64 queries), B = K MN-major
                mbar_wait(&dq_free[s], ((i >> 1) & 1) ^ 1);  // tile i-2 drained from this dQ buffer
                tc_fence_after_sync();
                constexpr uint32_t idesc_dq = make_idesc_bf16(64, HD, 1, 1);
                for (int k = 0; k < ksteps_kv; ++k)
                    umma_bf16(tmem_dq + s * 64, make_smem_desc_sw128(dst_addr + k * 2048, 128 * 128, 1024),
                              make_smem_desc_sw128(k_addr + k * 2048, 128 * 128, 1024), idesc_dq, k > 0 ? 1u : 0u);
                // (the S^T / dP^T MMAs of this tile, issued by the other thread, completed before pds_full fired)
                umma_commit(&pds_free[s]);
                umma_commit(&qdo_empty[r]);
                umma_commit(&dq_full[s]);
            }
            umma_commit(dkv_full);
        }
    } else if (warp >= ABF_E_WARPS) {
        // ===================== dQ drain warps =====================
        // UMMA M = 64 accumulator layout: row r lives in TMEM lane (r % 16) + 32 * (r / 16)
        const int quarter = warp & 3;
        const uint32_t lane_off = uint32_t(quarter * 32) << 16;
        for (int i = 0; i < nq; ++i) {
            const int s = i & 1;
            mbar_wait(&dq_full[s], (i >> 1) & 1);
            tc_fence_after_sync();
            const int q = i * 64 + quarter * 16 + lane;
            const bool ok = lane < 16 && q < a.N;
            float* dst = a.dq_acc + ((long long)b * a.N + q) * a.D + h * HD;
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_dq + s * 64 + lane_off + c * 16, r);
                tmem_ld_wait();
                if (ok) {
                    const float* f = reinterpret_cast<const float*>(r);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        red_add_v4_f32(dst + c * 16 + u * 4, f[4 * u], f[4 * u + 1], f[4 * u + 2], f[4 * u + 3]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&dq_free[s]);
        }
    } else {
        // ===================== elementwise warps =====================
        // two groups of 4 warps ping-pong on alternate query tiles (group g owns TMEM / smem buffer g): the TMEM reads,
        // exp / dS math and smem writes of tile i+1 overlap those of tile i. Thread = key row, all 64 queries.
        const int quarter = warp & 3, grp = warp >> 2;
        const int row = quarter * 32 + lane;          // key index within the block
        const uint32_t lane_off = uint32_t(quarter * 32) << 16;
        const int kv = kv0 + row;
        const bool row_ok = kv < a.N;
        const bool row_ok_warp = (kv0 + quarter * 32 + 31) < a.N;
        const int sw = row & 7;
        const long long stat_base = ((long long)b * a.H + h) * a.N;
        // lse2 / delta of every query of this (batch, head): staged once
        const int Np = nq * 64;
        for (int qi = threadIdx.x; qi < Np; qi += ABF_E_WARPS * 32) {
            stat[qi] = qi < a.N ? a.lse2[stat_base + qi] : 0.f;
            stat[Np + qi] = qi < a.N ? a.delta[stat_base + qi] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int s = grp;
        uint8_t* pt_row = smem + ABF_SMEM_PT + s * ABF_T128 + row * 128;
        uint8_t* dst_row = smem + ABF_SMEM_DST + s * ABF_T128 + row * 128;
        for (int i = grp; i < nq; i += 2) {
            const int valid = min(64, a.N - i * 64);
            const float* st_l = stat + i * 64;        // lse2 of this tile's queries (broadcast reads below)
            const float* st_d = stat + Np + i * 64;   // delta
            mbar_wait(&st_full[s], (i >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t sr[32], dpr[32];
                tmem_ld_32x32b_x32(tmem_st + s * 64 + lane_off + c * 32, sr);
                tmem_ld_32x32b_x32(tmem_dpt + s * 64 + lane_off + c * 32, dpr);
                tmem_ld_wait();
                if (c == 1) {
                    tc_fence_before_sync();
                    mbar_arrive(&st_free[s]);  // S^T / dP^T of this tile now live in registers
                }
                float p[32], ds[32];
                if (row_ok_warp && c * 32 + 32 <= valid) {  // warp-uniform fast path
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float l2 = st_l[c * 32 + q], dl = st_d[c * 32 + q];
                        p[q] = ex2_approx(fmaf(__uint_as_float(sr[q]), a.scale_log2, -l2));
                        ds[q] = a.scale * p[q] * (__uint_as_float(dpr[q]) - dl);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float l2 = st_l[c * 32 + q], dl = st_d[c * 32 + q];
                        const float pv = ex2_approx(fmaf(__uint_as_float(sr[q]), a.scale_log2, -l2));
                        const bool ok = row_ok && (c * 32 + q) < valid;
                        p[q] = ok ? pv : 0.f;
                        ds[q] = ok ? a.scale * pv * (__uint_as_float(dpr[q]) - dl) : 0.f;
                    }
                }
                // (waited for as late as possible) the MMAs of tile i-2 have read this smem buffer
                if (c == 0) mbar_wait(&pds_free[s], ((i >> 1) & 1) ^ 1);
                abf_store_row_chunk(pt_row, sw, c, p);
                abf_store_row_chunk(dst_row, sw, c, ds);
            }
            fence_proxy_async_smem();
            mbar_arrive(&pds_full[s]);
        }
        const int c = grp;  // write-out split: group 0 -> dK, group 1 -> dV
        mbar_wait(dkv_full, 0);
        tc_fence_after_sync();
        // warps 0-3 write dK, warps 4-7 write dV (thread = key row, 64 columns)
        __nv_bfloat16* dst = a.dqkv + ((long long)b * a.N + kv) * (3LL * a.D) + (c == 0 ? 1 : 2) * a.D + h * HD;
        const uint32_t tsrc = (c == 0 ? tmem_dk : tmem_dv) + lane_off;
#pragma unroll
        for (int cc = 0; cc < HD / 16; ++cc) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tsrc + cc * 16, r);
            tmem_ld_wait();
            if (row_ok) {
                const float* f = reinterpret_cast<const float*>(r);
                st_v4(dst + cc * 16, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                pack_bf16(f[6], f[7])));
                st_v4(dst + cc * 16 + 8, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]),
                                                    pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == ABF_W_MMA) {
        tc_fence_after_sync();
        tmem_dealloc<ABF_TMEM_COLS>(tmem_base);
    }
}

// delta[b,h,n] = sum_e dO[b,n,h,e] * O[b,n,h,e]      (thread = (token row, head); d = 64)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                  int B, int N, int H) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)B * N * H;
    if (t >= total) return;
    const int h = static_cast<int>(t % H);
    const long long row = t / H;  // b * N + n
    const __nv_bfloat16* op = o + row * (long long)H * 64 + h * 64;
    const __nv_bfloat16* dp = dout + row * (long long)H * 64 + h * 64;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const uint4 x = ld_nc_v4(op + u * 8);
        const uint4 y = ld_nc_v4(dp + u * 8);
        acc += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x);
        acc += bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y);
        acc += bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z);
        acc += bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
    }
    const long long bb = row / N;
    const int n = static_cast<int>(row - bb * N);
    delta[(bb * H + h) * N + n] = acc;
}

// dqkv[row, 0:D] = bf16(dq_acc[row, 0:D])   (row pitch of dqkv = 3D)
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int D) {
    const long long per_row = D / 8;
    const long long total = rows * per_row;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const long long row = t / per_row;
        const int c = static_cast<int>(t - row * per_row) * 8;
        const float4 x = *reinterpret_cast<const float4*>(acc + row * D + c);
        const float4 y = *reinterpret_cast<const float4*>(acc + row * D + c + 4);
        st_v4(dqkv + row * 3LL * D + c,
              make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w)));
    }
}

int make_tok_tmap2d(CUtensorMap* out, const void* p, long long rows, long long cols, int box_rows);  // attn_fwd.cu

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_attn_bwd_fused(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                                   float* delta, float* dq_f32_ws, void* dqkv_bf16, int B, int N, int H, int d,
                                   float scale, void* stream) {
    if (B <= 0 || N <= 0 || H <= 0 || d != 64) return VITK_ERR_ARG;
    if (!qkv_bf16 || !out_bf16 || !dout_bf16 || !lse2 || !delta || !dq_f32_ws || !dqkv_bf16) return VITK_ERR_ARG;
    const int nq_tiles = (N + 63) / 64;
    if (2 * nq_tiles * 64 * 4 > ABF_MAX_STAT_BYTES) return VITK_ERR_UNSUPPORTED;  // caller falls back to vitk_attn_bwd
    const int smem_bytes = abf_smem_bytes(nq_tiles);
    CUtensorMap q128, q64, do64;
    const long long rows = (long long)B * N;
    const int D = H * d;
    if (make_tok_tmap2d(&q128, qkv_bf16, rows, 3LL * D, 128) || make_tok_tmap2d(&q64, qkv_bf16, rows, 3LL * D, 64) ||
        make_tok_tmap2d(&do64, dout_bf16, rows, (long long)D, 64))
        return VITK_ERR_TMAP;
    static int attr_smem = 0;
    if (attr_smem < smem_bytes) {
        if (cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
            cudaSuccess)
            return VITK_ERR_CUDA;
        attr_smem = smem_bytes;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AttnBwdFusedArgs a;
    a.B = B; a.H = H; a.N = N; a.D = D;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.lse2 = lse2;
    a.delta = delta;
    a.dq_acc = dq_f32_ws;
    a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
    const long long nth = rows * H;
    attn_delta_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(out_bf16),
                                                                    reinterpret_cast<const __nv_bfloat16*>(dout_bf16),
                                                                    delta, B, N, H);
    if (cudaMemsetAsync(dq_f32_ws, 0, (size_t)rows * D * sizeof(float), st) != cudaSuccess) return VITK_ERR_CUDA;
    dim3 grid((N + 127) / 128, H, B);
    attn_bwd_fused_kernel<<<grid, ABF_THREADS, smem_bytes, st>>>(q128, q64, do64, a);
    if (cudaGetLastError() != cudaSuccess) return VITK_ERR_CUDA;
    long long cblocks = (rows * (D / 8) + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (cblocks > cap) cblocks = cap;
    attn_dq_convert_kernel<<<(unsigned)cblocks, 256, 0, st>>>(dq_f32_ws, a.dqkv, rows, D);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
