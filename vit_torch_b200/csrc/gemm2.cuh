// CTA-pair (tcgen05 cta_group::2) variant of the persistent bf16 GEMM for large problems.
//
//   D[M,N] = A[M,K] * B[N,K]^T, output tile 256 x 256 per 2-CTA cluster. Each CTA of the pair stages its own 128 rows of
//   A and HALF of the B tile (128 of the 256 B rows): per k-block an SM takes in 32 KB for 128x256x64 MACs instead of
//   the 48 KB of the 1-CTA kernel. The 1-CTA kernel is bound by exactly that L2 -> SM operand stream (ncu: tensor pipe
//   <= 70 % with every epilogue load queued behind the TMA traffic), so this is where the remaining GEMM time is.
//   One UMMA of M = 256 is issued by the leader CTA and executes on both SMs; each CTA's TMEM receives the 128 x 256
//   accumulator of its own rows, which its own epilogue warps drain (same EpilogueWarp engine as the 1-CTA kernel).
//
// Synchronisation (per CTA unless noted):
//   full[s]   (leader only) TMA bytes of BOTH CTAs for stage s        -> leader MMA thread
//   empty[s]  multicast tcgen05.commit, arrives in both CTAs          -> each CTA's TMA producer
//   tfull[a]  multicast tcgen05.commit, accumulator stage a complete  -> each CTA's epilogue warps
//   tempty[a] (leader only) 2 x 8 epilogue warps arrived (remote)     -> leader MMA thread
#pragma once
#include "gemm.cuh"

namespace vitk {

template <int EPI> struct Gemm2Cfg {
    static constexpr int BN = 256;
    static constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB: this CTA's 128 rows of A
    static constexpr int B_STAGE_BYTES = (BN / 2) * GEMM_BK * 2;       // 16 KB: this CTA's half of the B tile
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // 32 KB
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 256;
    static constexpr int VEC_BYTES = GEMM_EPI_WARPS * 2 * (BN / 2) * 4;
    static constexpr int STG_BYTES = EpiSmem<EPI>::STG_BYTES;           // TMA-store staging
    static constexpr int FIXED_BYTES = BAR_BYTES + VEC_BYTES + STG_BYTES + 2048;
    static constexpr int FIT_STAGES = (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT_STAGES < 6 ? FIT_STAGES : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
};

template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const GemmArgs g) {
    using Cfg = Gemm2Cfg<EPI>;
    constexpr int BN = Cfg::BN;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024 B alignment; both CTAs of the pair compute the same offsets (the UMMA
    // descriptors built by the leader address the peer's shared memory at identical offsets)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* svec = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
    uint8_t* sstg = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Cfg::VEC_BYTES) + 1023) &
        ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // unit = (m_unit of 256 rows, n_tile, split); this CTA owns rows [m_unit*256 + rank*128, +128)
    const int num_m_units = (g.num_m_tiles + 1) >> 1;
    const int num_units = num_m_units * g.num_n_tiles * g.splits;
    auto decode = [&](int un, int& m_unit, int& n_tile, int& split) {
        n_tile = un % g.num_n_tiles;
        const int rest = un / g.num_n_tiles;
        split = rest % g.splits;
        m_unit = rest / g.splits;
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                int m_unit, n_tile, split;
                decode(u, m_unit, n_tile, split);
                const int m0 = m_unit * (2 * GEMM_BM) + rank * GEMM_BM;
                const int n0 = n_tile * BN + rank * (BN / 2);
                const int kb0 = split * g.kblocks_per_split;
                const int kb1 = min(kb0 + g.kblocks_per_split, g.num_kblocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait_role(&empty_bar[stage], phase ^ 1, g.dbg);
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    uint8_t* sa = smem_a + stage * Cfg::A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
                    const int k0 = kb * GEMM_BK;
                    if constexpr (!A_MN) {
                        tma_load_2d_2cta(sa, &tmA, &full_bar[stage], k0, m0);  // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j)  // box {64 m, 64 k} per 64-wide M chunk
                            tma_load_2d_2cta(sa + j * (GEMM_BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, k0);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d_2cta(sb, &tmB, &full_bar[stage], k0, n0);  // box {64 k, 128 n}
                    } else {
#pragma unroll
                        for (int j = 0; j < (BN / 2) / 64; ++j)
                            tma_load_2d_2cta(sb + j * (GEMM_BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, k0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            constexpr uint32_t A_LBO = A_MN ? GEMM_BK * 128 : 0, B_LBO = B_MN ? GEMM_BK * 128 : 0;
            constexpr uint32_t A_KSTEP = (A_MN ? 2048u : 32u) >> 4, B_KSTEP = (B_MN ? 2048u : 32u) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                int m_unit, n_tile, split;
                decode(u, m_unit, n_tile, split);
                const int kb0 = split * g.kblocks_per_split;
                const int kb1 = min(kb0 + g.kblocks_per_split, g.num_kblocks);
                mbar_wait_role(&tempty_bar[as], aphase ^ 1, g.dbg);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint64_t adesc =
                        make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::A_STAGE_BYTES), A_LBO, 1024);
                    const uint64_t bdesc =
                        make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES), B_LBO, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_bf16_2cta(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc,
                                       (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit_2cta(&empty_bar[stage]);  // frees this smem stage in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&tfull_bar[as]);  // accumulator complete -> epilogue warps of both CTAs
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= GEMM_EPI_WARP0) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int ew = warp - GEMM_EPI_WARP0;
        EpilogueWarp<BN, EPI> epi(g, svec, sstg, &tmO, &tmO2, ew, warp, lane);
        int as = 0;
        uint32_t aphase = 0;
        int m_unit, n_tile, split;
        if (cluster_id < num_units) {
            decode(cluster_id, m_unit, n_tile, split);
            epi.l2_prefetch(2 * m_unit + rank, n_tile, 0);
        }
        for (int u = cluster_id; u < num_units; u += num_clusters) {
            if (u + num_clusters < num_units) {  // a whole tile of lead time
                decode(u + num_clusters, m_unit, n_tile, split);
                epi.l2_prefetch(2 * m_unit + rank, n_tile, 0);
            }
            decode(u, m_unit, n_tile, split);
            const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + as * BN + (ew >> 2) * (BN / 2);
            epi.tile(2 * m_unit + rank, n_tile, 0, taddr, [&]() { mbar_wait(&tfull_bar[as], aphase); },
                     [&]() { if (lane == 0) mbar_arrive_leader(&tempty_bar[as]); });
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        epi.finish();
    }

    // no CTA may exit (or free TMEM) while its peer can still touch its shared memory, barriers or TMEM
    tc_fence_before_sync();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace vitk
