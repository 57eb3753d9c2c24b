// Whole-head attention backward for short sequences (N <= 256: ViT-S/B 16 at 224 px, DeiT, 196..198 tokens):
// ONE thread block owns one (image, head) and produces dQ, dK and dV in a single pass -- S and dP are computed once
// (the two-kernel backward of attn_bwd.cu computes them twice: 7 tile-GEMMs and two elementwise passes per tile pair),
// nothing is accumulated through global memory, delta = rowsum(dO o O) is formed in the prologue.
//
//   P = exp2(S*scale*log2e - lse2),  dP = dO V^T,  dS = scale * P o (dP - delta)
//   dV = P^T dO        dK = dS^T Q        dQ = dS K                                  
//
// All of Q, K, V, dO of the head (<= 256 rows x 64) are staged once by TMA (128 KB of shared memory). For each block of
// 128 keys (kt) and each tile of 64 queries (i):
//     S^T = K_kt Q_i^T, dP^T = V_kt dO_i^T     UMMA 128 x 64 x 16 into 64 + 64 TMEM columns
//     8 elementwise warps (thread = key row, group g = 32 of the 64 queries) form P^T and dS^T as bf16 in swizzled smem
//     dV_kt += P^T dO_i,  dK_kt += dS^T Q_i    UMMA 128 x 64, B read MN-major
//     dQ_i  += dS K_kt                         UMMA M = 64: A is the SAME dS^T tile read MN-major, B = K_kt MN-major;
//                                              the 4 query tiles' accumulators stay in TMEM across both key blocks
// TMEM: 64 (S^T) + 64 (dP^T) + 64 (dV) + 64 (dK) + 4 x 64 (dQ) = 512 columns -> one block per SM.
// Outputs leave as bf16 into dqkv [B, N, 3, H, d] (dK / dV through swizzled staging tiles + one rank-3 TMA store per key
// block; dQ by 16-byte stores); the column sums of dQ / dK / dV (= the qkv bias gradient) are reduced on the way out.
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int ABH_THREADS = 320;          // warps 0-7 elementwise, warp 8 TMA, warp 9 MMA
constexpr int ABH_T128 = 128 * 128;       // [128 rows x 64 bf16] swizzled tile
constexpr int ABH_T64 = 64 * 128;         // [64 rows x 64 bf16]
constexpr int ABH_MAX_N = 256;
constexpr int ABH_SMEM_K = 0;                              // 2 tiles [128 x 64]
constexpr int ABH_SMEM_V = ABH_SMEM_K + 2 * ABH_T128;      // 2 tiles
constexpr int ABH_SMEM_Q = ABH_SMEM_V + 2 * ABH_T128;      // 4 tiles [64 x 64]
constexpr int ABH_SMEM_DO = ABH_SMEM_Q + 4 * ABH_T64;      // 4 tiles
constexpr int ABH_SMEM_PT = ABH_SMEM_DO + 4 * ABH_T64;     // P^T  [128 kv x 64 q]
constexpr int ABH_SMEM_DST = ABH_SMEM_PT + ABH_T128;       // dS^T [128 kv x 64 q]
constexpr int ABH_SMEM_STAT = ABH_SMEM_DST + ABH_T128;     // lse2[256] | delta[256] fp32
constexpr int ABH_SMEM_BAR = ABH_SMEM_STAT + 2 * ABH_MAX_N * 4;
constexpr int ABH_SMEM_BYTES = ABH_SMEM_BAR + 256;
constexpr uint32_t ABH_TMEM_COLS = 512;

struct AttnBwdHeadArgs {
    int B, H, N, D;
    float scale, scale_log2;
    const __nv_bfloat16* out;   // forward output O [B*N, D]
    const __nv_bfloat16* dout;  // dO [B*N, D]
    const float* lse2;          // [B,H,N]
    __nv_bfloat16* dqkv;        // [B*N, 3D]
    float* dbias;               // optional fp32 [3D], += column sums of dqkv
};

__device__ __forceinline__ void abh_colsum_chunk(float* dst16, const uint32_t* r, bool row_ok, int lane) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = row_ok ? __uint_as_float(r[i]) : 0.f;
    const float tot = warp_colsum16(v, lane);
    if ((lane & 1) == 0) atomicAdd(dst16 + warp_colsum16_col(lane), tot);
}
__device__ __forceinline__ void abh_store_unit(uint8_t* tile_row, int sw, int unit, const float* v) {
    *reinterpret_cast<uint4*>(tile_row + ((unit ^ sw) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

__global__ void __launch_bounds__(ABH_THREADS, 1)
attn_bwd_head_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                     const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmDQKV,
                     const AttnBwdHeadArgs a) {
    constexpr int HD = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ABH_SMEM_BAR);
    uint64_t* kv_full = bars + 0;     // [2]
    uint64_t* qdo_full = bars + 2;    // [4]
    uint64_t* st_full = bars + 6;     // S^T / dP^T ready in TMEM
    uint64_t* st_free = bars + 7;
    uint64_t* pds_full = bars + 8;    // P^T / dS^T written to smem
    uint64_t* pds_free = bars + 9;
    uint64_t* dkv_full = bars + 10;   // dV / dK of a key block complete
    uint64_t* dkv_free = bars + 11;   // ... and read out of TMEM
    uint64_t* dq_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* stat_l = reinterpret_cast<float*>(smem + ABH_SMEM_STAT);
    float* stat_d = stat_l + ABH_MAX_N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int nq = (a.N + 63) / 64, nkt = (a.N + 127) / 128;
    const int nit = nq * nkt;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmQKV128);
        tma_prefetch_desc(&tmQKV64);
        tma_prefetch_desc(&tmDO64);
        for (int i = 0; i < 2; ++i) mbar_init(&kv_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&qdo_full[i], 1);
        mbar_init(st_full, 1);
        mbar_init(st_free, 256);
        mbar_init(pds_full, 256);
        mbar_init(pds_free, 1);
        mbar_init(dkv_full, 1);
        mbar_init(dkv_free, 256);
        mbar_init(dq_full, 1);
        fence_mbar_init();
        // everything the head needs, in the order the MMA warp consumes it
        mbar_expect_tx(&kv_full[0], 2 * ABH_T128);
        tma_load_2d(smem + ABH_SMEM_K, &tmQKV128, &kv_full[0], (a.H + h) * HD, b * a.N);
        tma_load_2d(smem + ABH_SMEM_V, &tmQKV128, &kv_full[0], (2 * a.H + h) * HD, b * a.N);
        for (int i = 0; i < nq; ++i) {
            mbar_expect_tx(&qdo_full[i], 2 * ABH_T64);
            tma_load_2d(smem + ABH_SMEM_Q + i * ABH_T64, &tmQKV64, &qdo_full[i], h * HD, b * a.N + i * 64);
            tma_load_2d(smem + ABH_SMEM_DO + i * ABH_T64, &tmDO64, &qdo_full[i], h * HD, b * a.N + i * 64);
        }
        if (nkt > 1) {
            mbar_expect_tx(&kv_full[1], 2 * ABH_T128);
            tma_load_2d(smem + ABH_SMEM_K + ABH_T128, &tmQKV128, &kv_full[1], (a.H + h) * HD, b * a.N + 128);
            tma_load_2d(smem + ABH_SMEM_V + ABH_T128, &tmQKV128, &kv_full[1], (2 * a.H + h) * HD, b * a.N + 128);
        }
    }
    if (warp == 9) tmem_alloc<ABH_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 64, tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 192,
                   tmem_dq = tmem_base + 256;

    if (warp == 8) {
        // (all loads were issued in the prologue)
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t pt_addr = smem_u32(smem + ABH_SMEM_PT), dst_addr = smem_u32(smem + ABH_SMEM_DST);
            auto issue_st = [&](int it) {
                const int kt = it / nq, i = it - kt * nq;
                const int valid = min(64, a.N - i * 64);
                const uint32_t idesc = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
                mbar_wait_backoff(&kv_full[kt], 0);
                mbar_wait_backoff(&qdo_full[i], 0);
                if (it > 0) mbar_wait_backoff(st_free, (it - 1) & 1);
                tc_fence_after_sync();
                const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + ABH_SMEM_K + kt * ABH_T128), 0, 1024);
                const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + ABH_SMEM_V + kt * ABH_T128), 0, 1024);
                const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + ABH_SMEM_Q + i * ABH_T64), 0, 1024);
                const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + ABH_SMEM_DO + i * ABH_T64), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_st, kd + 2 * k, qd + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_dpt, vd + 2 * k, dod + 2 * k, idesc, k > 0);
                umma_commit(st_full);
            };
            issue_st(0);
            for (int it = 0; it < nit; ++it) {
                const int kt = it / nq, i = it - kt * nq;
                if (it + 1 < nit) issue_st(it + 1);
                const int valid = min(64, a.N - i * 64);
                const int ksteps = (valid + 15) >> 4;
                const int kv_valid = min(128, a.N - kt * 128);
                const int ksteps_kv = (kv_valid + 15) >> 4;
                mbar_wait_backoff(pds_full, it & 1);
                if (i == 0 && kt > 0) mbar_wait_backoff(dkv_free, (kt - 1) & 1);   // previous block's dV / dK read out
                tc_fence_after_sync();
                constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t q_addr = smem_u32(smem + ABH_SMEM_Q + i * ABH_T64);
                const uint32_t do_addr = smem_u32(smem + ABH_SMEM_DO + i * ABH_T64);
                const uint32_t k_addr = smem_u32(smem + ABH_SMEM_K + kt * ABH_T128);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    umma_bf16(tmem_dv, make_smem_desc_sw128(pt_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(do_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                    umma_bf16(tmem_dk, make_smem_desc_sw128(dst_addr + k * 32, 0, 1024),
                              make_smem_desc_sw128(q_addr + k * 2048, 64 * 128, 1024), idesc_acc, acc);
                }
                // dQ_i[q, HD] += dS[q, kv] K_kt[kv, HD]: A = the dS^T tile read MN-major (M = 64 queries), B = K MN-major
                constexpr uint32_t idesc_dq = make_idesc_bf16(64, HD, 1, 1);
                for (int k = 0; k < ksteps_kv; ++k)
                    umma_bf16(tmem_dq + i * 64, make_smem_desc_sw128(dst_addr + k * 2048, 128 * 128, 1024),
                              make_smem_desc_sw128(k_addr + k * 2048, 128 * 128, 1024), idesc_dq,
                              (kt > 0 || k > 0) ? 1u : 0u);
                umma_commit(pds_free);
                if (i == nq - 1) umma_commit(dkv_full);
            }
            umma_commit(dq_full);
        }
    } else {
        // ===================== elementwise warps =====================
        const int g = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;                       // key row within a key block
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        uint8_t* pt_row = smem + ABH_SMEM_PT + row * 128;
        uint8_t* dst_row = smem + ABH_SMEM_DST + row * 128;
        const int sw = row & 7;
        // ---- prologue: lse2 and delta = rowsum(dO o O) of every query of the head (thread = query row)
        {
            const int q = threadIdx.x;
            float l2 = 0.f, dl = 0.f;
            if (q < a.N) {
                l2 = a.lse2[((long long)b * a.H + h) * a.N + q];
                const __nv_bfloat16* op = a.out + ((long long)b * a.N + q) * a.D + h * HD;
                const __nv_bfloat16* dp = a.dout + ((long long)b * a.N + q) * a.D + h * HD;
                uint4 x[8], y[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    x[u] = ld_nc_v4(op + u * 8);
                    y[u] = ld_nc_v4(dp + u * 8);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    dl += bf16_lo(x[u].x) * bf16_lo(y[u].x) + bf16_hi(x[u].x) * bf16_hi(y[u].x);
                    dl += bf16_lo(x[u].y) * bf16_lo(y[u].y) + bf16_hi(x[u].y) * bf16_hi(y[u].y);
                    dl += bf16_lo(x[u].z) * bf16_lo(y[u].z) + bf16_hi(x[u].z) * bf16_hi(y[u].z);
                    dl += bf16_lo(x[u].w) * bf16_lo(y[u].w) + bf16_hi(x[u].w) * bf16_hi(y[u].w);
                }
            }
            stat_l[q] = l2;
            stat_d[q] = dl;
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        for (int it = 0; it < nit; ++it) {
            const int kt = it / nq, i = it - kt * nq;
            const int kv = kt * 128 + row;
            const bool row_ok = kv < a.N;
            const bool warp_live = (kt * 128 + wq * 32) < a.N;          // warp-uniform
            const bool row_ok_warp = (kt * 128 + wq * 32 + 31) < a.N;
            const int valid = min(64, a.N - i * 64);
            const int nv = min(32, valid - 32 * g);
            const bool work = warp_live && nv > 0;                       // warp-uniform
            const float4* st4 = reinterpret_cast<const float4*>(stat_l + i * 64 + g * 32);
            const float4* dl4 = reinterpret_cast<const float4*>(stat_d + i * 64 + g * 32);
            mbar_wait(st_full, it & 1);
            tc_fence_after_sync();
            uint32_t sr[32], dpr[32];
            if (work) {
                tmem_ld_32x32b_x32(tmem_st + lane_off + g * 32, sr);
                tmem_ld_32x32b_x32(tmem_dpt + lane_off + g * 32, dpr);
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            mbar_arrive(st_free);                            // S^T / dP^T columns go back to the MMA warp before the math
            if (it > 0) mbar_wait(pds_free, (it - 1) & 1);   // previous dV / dK / dQ MMAs finished reading P^T / dS^T
            if (work) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float p[8], ds[8];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float4 l4 = st4[2 * u + hh], d4 = dl4[2 * u + hh];
                        const float l2[4] = {l4.x, l4.y, l4.z, l4.w};
                        const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
                        if (row_ok_warp && nv == 32) {
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
                    abh_store_unit(pt_row, sw, 4 * g + u, p);
                    abh_store_unit(dst_row, sw, 4 * g + u, ds);
                }
                fence_proxy_async_smem();
            } else if (nv > 0) {
                // rows of this warp lie past the end of the image: the dV / dK / dQ MMAs still read them -> zeros
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    *reinterpret_cast<uint4*>(pt_row + (((4 * g + u) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(dst_row + (((4 * g + u) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_proxy_async_smem();
            }
            mbar_arrive(pds_full);

            if (i == nq - 1) {
                // ---- key block kt complete: group 0 writes dK, group 1 writes dV
                mbar_wait(dkv_full, kt & 1);
                tc_fence_after_sync();
                uint8_t* stage_row = g ? dst_row : pt_row;   // (free: dkv_full covers every MMA that read them)
                const uint32_t tmem_acc = g ? tmem_dv : tmem_dk;
                float* dbias = a.dbias != nullptr ? a.dbias + (1 + g) * a.D + h * HD : nullptr;
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(tmem_acc + lane_off + c * 16, r);
                    tmem_ld_wait();
                    if (dbias != nullptr && warp_live) abh_colsum_chunk(dbias + c * 16, r, row_ok, lane);
                    const float* f = reinterpret_cast<const float*>(r);
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c) ^ sw) << 4)) =
                        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c + 1) ^ sw) << 4)) =
                        make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                   pack_bf16(f[14], f[15]));
                }
                tc_fence_before_sync();
                mbar_arrive(dkv_free);
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (threadIdx.x == 0) {
                    tma_store_3d(&tmDQKV, smem + ABH_SMEM_PT, a.D + h * HD, kt * 128, b);        // dK (rows >= N clipped)
                    tma_store_3d(&tmDQKV, smem + ABH_SMEM_DST, 2 * a.D + h * HD, kt * 128, b);   // dV
                    tma_store_commit();
                    tma_store_wait_read<0>();      // the staging tiles are rewritten by the next key block
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
        // ---- dQ: UMMA M = 64 accumulator layout, row r of tile i lives in TMEM lane (r % 16) + 32 * (r / 16)
        mbar_wait(dq_full, 0);
        tc_fence_after_sync();
        for (int i = g; i < nq; i += 2) {
            const int q = i * 64 + wq * 16 + lane;
            const bool ok = lane < 16 && q < a.N;
            __nv_bfloat16* dst = a.dqkv + ((long long)b * a.N + q) * (3LL * a.D) + h * HD;
            float* dbias = a.dbias != nullptr ? a.dbias + h * HD : nullptr;
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_dq + i * 64 + lane_off + c * 16, r);
                tmem_ld_wait();
                if (dbias != nullptr) abh_colsum_chunk(dbias + c * 16, r, ok, lane);
                if (ok) {
                    const float* f = reinterpret_cast<const float*>(r);
                    st_v4(dst + c * 16, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                   pack_bf16(f[6], f[7])));
                    st_v4(dst + c * 16 + 8, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]),
                                                       pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
                }
            }
        }
        if (threadIdx.x == 0) tma_store_wait<0>();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<ABH_TMEM_COLS>(tmem_base);
    }
}

int make_tok_tmap2d(CUtensorMap* out, const void* p, long long rows, long long cols, int box_rows);  // attn_fwd.cu

}  // namespace vitk

using namespace vitk;

// Opt-in (VITK_ATTN_BWD_HEAD=1): measured on B200, ViT-B/16 bs128: 281 us vs 222 us for the two-kernel backward. With
// 512 TMEM columns the block is alone on its SM and its per-tile chain (S^T MMA -> exp / dS -> dV / dK / dQ MMAs) is
// latency-bound, so the SM works on ONE tile at a time where two co-resident blocks of the two-kernel form overlap two;
// the saved recompute does not make up for that. Kept parity-tested for the next step (two tiles in flight per block).
extern "C" int vitk_attn_bwd_head_supported(int N, int d) {
    const char* e = getenv("VITK_ATTN_BWD_HEAD");
    if (e == nullptr || e[0] != '1') return 0;
    return (d == 64 && N >= 1 && N <= ABH_MAX_N) ? 1 : 0;
}

extern "C" int vitk_attn_bwd_head(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                                  void* dqkv_bf16, float* dqkv_bias_grad, int B, int N, int H, int d, float scale,
                                  void* stream) {
    if (B <= 0 || H <= 0 || !qkv_bf16 || !out_bf16 || !dout_bf16 || !lse2 || !dqkv_bf16) return VITK_ERR_ARG;
    if (d != 64 || N < 1 || N > ABH_MAX_N) return VITK_ERR_UNSUPPORTED;
    CUtensorMap q128, q64, do64, dqkv3;
    const long long rows = (long long)B * N;
    const int D = H * d;
    if (make_tmap_3d_tok_store(&dqkv3, dqkv_bf16, 3ull * D, (uint64_t)N, (uint64_t)B, 128)) return VITK_ERR_TMAP;
    if (make_tok_tmap2d(&q128, qkv_bf16, rows, 3LL * D, 128) || make_tok_tmap2d(&q64, qkv_bf16, rows, 3LL * D, 64) ||
        make_tok_tmap2d(&do64, dout_bf16, rows, (long long)D, 64))
        return VITK_ERR_TMAP;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABH_SMEM_BYTES) !=
            cudaSuccess)
            return VITK_ERR_CUDA;
        attr = true;
    }
    AttnBwdHeadArgs a;
    a.B = B; a.H = H; a.N = N; a.D = D;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out = reinterpret_cast<const __nv_bfloat16*>(out_bf16);
    a.dout = reinterpret_cast<const __nv_bfloat16*>(dout_bf16);
    a.lse2 = lse2;
    a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
    a.dbias = dqkv_bias_grad;
    attn_bwd_head_kernel<<<dim3(H, B), ABH_THREADS, ABH_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
        q128, q64, do64, dqkv3, a);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
