// Fused flash-style attention forward on tcgen05:  O = softmax(Q K^T * scale) V   per (batch, head).
//
// Replaces on the reference path: timm/DINO Attention.forward core (in-repo witness
// models/swin.py:119-144): attn = (q @ k^T) * scale; softmax; attn @ v -- without materialising [B,H,N,N].
//
// Layout: qkv is the raw output of the qkv Linear, bf16 [B, N, 3, H, d] (= rows [B*N, 3*H*d]); Q/K/V tiles are
// fetched straight out of it with 2-D TMA maps over [B*N, 3*H*d] (128B swizzle; rows >= N of a tile belong to the next
// image and are masked, rows past the tensor end are zero-filled by TMA). O is written bf16 [B, N, H, d] (heads
// merged = the layout proj consumes), lse2[b,h,n] = log2-domain logsumexp of the scaled scores (saved for backward).
//
// One CTA = one (q-block of 128 rows, head, batch). Warps 0-3: softmax (thread == query row, TMEM lane == row),
// warp 4: TMA producer, warp 5: MMA issuer + TMEM allocator. Key/value tiles are 64 rows.
//
// The kernel is bound by TMEM read bandwidth (64 B/clk/SM) and instruction issue, not by the tensor pipe, so every
// score is read from TMEM exactly ONCE and the output accumulator never leaves TMEM:
//   * S_j = Q K_j^T (UMMA 128 x 64 x 16 into 64 TMEM columns) is loaded into registers in one go, which frees the S
//     columns for the next tile immediately;
//   * exponentials use a per-row reference maximum m_ref that is only raised (and O / l rescaled, through
//     tcgen05.ld/st) when a row maximum exceeds it by more than 2^8 (lazy rescaling: p <= 256 keeps fp32/bf16 exact
//     enough, lse2 = m_ref + log2(l) stays exact);
//   * P_j (bf16, 128B-swizzled K-major smem tile, double-buffered) x V_j (MN-major B operand read from the token-major
//     tile) accumulates straight into the 64 O columns of TMEM; O is read once at the end.
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int AF_BQ = 128;
constexpr int AF_BKV = 64;
constexpr int AF_THREADS = 192;
constexpr int AF_QTILE = 128 * 128;                       // [128 rows x 64 bf16] swizzled tile
constexpr int AF_KVTILE = AF_BKV * 128;                   // [64 rows x 64 bf16]
constexpr int AF_SMEM_Q = 0;
constexpr int AF_SMEM_K = AF_SMEM_Q + AF_QTILE;           // 2 stages
constexpr int AF_SMEM_V = AF_SMEM_K + 2 * AF_KVTILE;      // 2 stages
constexpr int AF_SMEM_P = AF_SMEM_V + 2 * AF_KVTILE;      // 2 buffers [128 x 64] bf16
constexpr int AF_SMEM_BAR = AF_SMEM_P + 2 * AF_QTILE;     // 81920
constexpr int AF_SMEM_BYTES = AF_SMEM_BAR + 256;
constexpr uint32_t AF_TMEM_COLS = 128;                    // S: [0,64)  O: [64,128)
constexpr float AF_RESCALE_TAU = 8.0f;                    // log2 units

struct AttnFwdArgs {
    int B, H, N, D;  // D = H * d
    float scale_log2;  // scale * log2(e)
    __nv_bfloat16* out;  // [B*N, D]
    float* lse2;         // [B, H, N]
    long long* trace;    // instrumented build only (VITK_TRACE), else nullptr
    int pf_dist;         // two-group kernel: CTA i pulls the tiles of CTA i + pf_dist into L2 (0 = off)
    int mma_spin;        // two-group kernel: the MMA warp polls its barriers without back-off (VITK_ATTN_MMA_SPIN=1)
};

#ifdef VITK_TRACE
long long* g_attn_trace = nullptr;
#endif

template <int HD>
__global__ void __launch_bounds__(AF_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const AttnFwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* v_full = bars + 3;    // [2]
    uint64_t* kv_empty = bars + 5;  // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_full = bars + 9;    // [2]
    uint64_t* o_full = bars + 11;   // [2]  PV_j complete (j even / odd)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qblk * AF_BQ;
    const int nkv = (a.N + AF_BKV - 1) / AF_BKV;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1 KB alignment
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        fence_mbar_init();
        // first loads go out before the block-wide sync: their latency overlaps the TMEM allocation
        mbar_expect_tx(q_full, AF_QTILE);
        tma_load_2d(smem + AF_SMEM_Q, &tmQ, q_full, h * HD, b * a.N + q0);
        for (int j = 0; j < 2 && j < nkv; ++j) {
            mbar_expect_tx(&k_full[j], AF_KVTILE);
            tma_load_2d(smem + AF_SMEM_K + j * AF_KVTILE, &tmKV, &k_full[j], (a.H + h) * HD, b * a.N + j * AF_BKV);
            mbar_expect_tx(&v_full[j], AF_KVTILE);
            tma_load_2d(smem + AF_SMEM_V + j * AF_KVTILE, &tmKV, &v_full[j], (2 * a.H + h) * HD, b * a.N + j * AF_BKV);
        }
    }
    if (warp == 5) tmem_alloc<AF_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + AF_BKV;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int j = 2; j < nkv; ++j) {   // tiles 0 and 1 were issued in the prologue
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], AF_KVTILE);
                tma_load_2d(smem + AF_SMEM_K + s * AF_KVTILE, &tmKV, &k_full[s], (a.H + h) * HD, b * a.N + j * AF_BKV);
                mbar_expect_tx(&v_full[s], AF_KVTILE);
                tma_load_2d(smem + AF_SMEM_V + s * AF_KVTILE, &tmKV, &v_full[s], (2 * a.H + h) * HD,
                            b * a.N + j * AF_BKV);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + AF_SMEM_Q), p_addr = smem_u32(smem + AF_SMEM_P);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                const int valid = min(AF_BKV, a.N - j * AF_BKV);
                const int ncols = (valid + 15) & ~15;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                if (j > 0) mbar_wait(s_free, (j - 1) & 1);
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_bf16(128, ncols, 0, 0);
                const uint64_t adesc = make_smem_desc_sw128(q_addr, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + AF_SMEM_K + s * AF_KVTILE), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                if (j + 1 < nkv) issue_s(j + 1);
                const int valid = min(AF_BKV, a.N - j * AF_BKV);
                const int ksteps = (valid + 15) >> 4;
                mbar_wait(&p_full[s], (j >> 1) & 1);
                mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_fence_after_sync();
                // O[128, HD] += P_j[128, kv] * V_j[kv, HD]: A = P (K-major), B = V tile read MN-major
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t v_addr = smem_u32(smem + AF_SMEM_V + s * AF_KVTILE);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t adesc = make_smem_desc_sw128(p_addr + s * AF_QTILE + k * 32, 0, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(v_addr + k * 2048, AF_BKV * 128, 1024);
                    umma_bf16(tmem_o, adesc, bdesc, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&o_full[s]);
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        // ===================== softmax (thread == query row) =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        float m_ref = -INFINITY, l_run = 0.f;
        const int sw = row & 7;

        for (int j = 0; j < nkv; ++j) {
            const int valid = min(AF_BKV, a.N - j * AF_BKV);
            mbar_wait(s_full, j & 1);
            tc_fence_after_sync();
            // the whole score tile goes to registers in one TMEM pass; the S columns are free again right away
            uint32_t sr[AF_BKV];
            tmem_ld_32x32b_x32(tmem_s + lane_off, sr);
            if (valid > 32) tmem_ld_32x32b_x32(tmem_s + lane_off + 32, sr + 32);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(s_free);
            // four independent chains: with one or two warps per scheduler a 64-deep dependent FMNMX / FADD chain is
            // pure latency
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (valid == AF_BKV) {
#pragma unroll
                for (int i = 0; i < AF_BKV; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sr[i]));
            } else {
#pragma unroll
                for (int i = 0; i < AF_BKV; ++i)
                    mx4[i & 3] = fmaxf(mx4[i & 3], (i < valid) ? __uint_as_float(sr[i]) : -INFINITY);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const float mxs = mx * a.scale_log2;
            if (j == 0) {
                m_ref = mxs;
            } else if (__any_sync(0xffffffffu, mxs > m_ref + AF_RESCALE_TAU)) {
                // lazy rescale (warp-uniform branch): raise the reference maximum and rescale l and the TMEM accumulator
                mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // PV_{j-1} has landed in O
                tc_fence_after_sync();
                const float m_new = fmaxf(m_ref, mxs);
                const float alpha = ex2_approx(m_ref - m_new);
                l_run *= alpha;
                m_ref = m_new;
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(tmem_o + lane_off + c * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32b_x16(tmem_o + lane_off + c * 16, r);
                }
                tmem_st_wait();
                tc_fence_before_sync();
            }
            // P buffer (j & 1) was last read by PV_{j-2}
            if (j >= 2) mbar_wait(&o_full[j & 1], ((j - 2) >> 1) & 1);
            uint8_t* p_row = smem + AF_SMEM_P + (j & 1) * AF_QTILE + row * 128;
            float ps4[4] = {0.f, 0.f, 0.f, 0.f};
            const float nm = -m_ref;
            if (valid == AF_BKV) {
#pragma unroll
                for (int u = 0; u < AF_BKV / 8; ++u) {
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        p[i] = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, nm));
                        ps4[i & 3] += p[i];
                    }
                    *reinterpret_cast<uint4*>(p_row + ((u ^ sw) << 4)) = make_uint4(
                        pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                }
            } else {
#pragma unroll
                for (int u = 0; u < AF_BKV / 8; ++u) {
                    if (u < ((valid + 15) >> 4) * 2) {   // 16-byte units of 8 columns that the PV MMA may read
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float v = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, nm));
                            p[i] = (u * 8 + i < valid) ? v : 0.f;
                            ps4[i & 3] += p[i];
                        }
                        *reinterpret_cast<uint4*>(p_row + ((u ^ sw) << 4)) = make_uint4(
                            pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                    }
                }
            }
            const float psum = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
            l_run += psum;
            fence_proxy_async_smem();
            mbar_arrive(&p_full[j & 1]);
        }
        // epilogue: the accumulator is read from TMEM exactly once
        mbar_wait(&o_full[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
        tc_fence_after_sync();
        const int n = q0 + row;
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = a.out + ((long long)b * a.N + n) * a.D + h * HD;
        uint32_t r[HD];
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) tmem_ld_32x32b_x16(tmem_o + lane_off + c * 16, r + c * 16);  // warp-collective
        tmem_ld_wait();
        if (n < a.N) {
            const float* f = reinterpret_cast<const float*>(r);
#pragma unroll
            for (int u = 0; u < HD / 8; ++u)
                st_v4(dst + u * 8, make_uint4(pack_bf16(f[u * 8 + 0] * inv, f[u * 8 + 1] * inv),
                                              pack_bf16(f[u * 8 + 2] * inv, f[u * 8 + 3] * inv),
                                              pack_bf16(f[u * 8 + 4] * inv, f[u * 8 + 5] * inv),
                                              pack_bf16(f[u * 8 + 6] * inv, f[u * 8 + 7] * inv)));
        }
        if (n < a.N) a.lse2[((long long)b * a.H + h) * a.N + n] = m_ref + log2f(l_run);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc<AF_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Two-group variant (default): the same tile pipeline with EIGHT softmax warps per CTA. Group g = warp / 4 owns the key
// columns [32g, 32g + 32) of every 64-key tile and runs its own online softmax over them (its own reference maximum,
// running sum and output accumulator O_g in TMEM: PV k-steps 2g, 2g+1 of a tile accumulate into O_g), so the two groups
// never exchange anything inside the loop; the epilogue merges (m_0, l_0, O_0) and (m_1, l_1, O_1) per row. A thread
// handles 32 scores per tile instead of 64: the dependent LDTM -> max -> exp -> pack -> STS chain that bounds the
// per-tile latency is half as long and twice as many warps are there to cover it. Warps whose 32 query rows all lie
// past the end of the image only take part in the barriers.
constexpr int AF2_THREADS = 320;                          // warps 0-7 softmax, warp 8 TMA, warp 9 MMA
constexpr int AF2_NS = 4;                                 // K/V ring stages: a 197-token image is fully in flight at once
constexpr int AF2_SMEM_Q = 0;                             // (reused for the groups' (m_ref, l) exchange in the epilogue)
constexpr int AF2_SMEM_K = AF2_SMEM_Q + AF_QTILE;
constexpr int AF2_SMEM_V = AF2_SMEM_K + AF2_NS * AF_KVTILE;
constexpr int AF2_SMEM_P = AF2_SMEM_V + AF2_NS * AF_KVTILE;   // 2 buffers [128 x 64] bf16
constexpr int AF2_SMEM_BAR = AF2_SMEM_P + 2 * AF_QTILE;
constexpr int AF2_SMEM_BYTES = AF2_SMEM_BAR + 256;
constexpr uint32_t AF2_TMEM_COLS = 256;                   // S: [0,64)  O_0: [64,128)  O_1: [128,192)

template <int HD>
__global__ void __launch_bounds__(AF2_THREADS, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnFwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NS = AF2_NS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF2_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* s_full = bars + 1;
    uint64_t* s_free = bars + 2;
    uint64_t* p_full = bars + 3;    // [group][buffer]
    uint64_t* o_full = bars + 7;    // [2]  PV_j complete (j even / odd), both groups
    uint64_t* k_full = bars + 9;    // [NS]
    uint64_t* v_full = k_full + NS;
    uint64_t* kv_empty = v_full + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qblk * AF_BQ;
    const int nkv = (a.N + AF_BKV - 1) / AF_BKV;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1 KB alignment
    if (warp == 8 && lane == 0) {
        VITK_TRACE_EV(a.trace, 0);
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&p_full[i], 128);
            mbar_init(&p_full[2 + i], 128);
            mbar_init(&o_full[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 256);
        fence_mbar_init();
        mbar_expect_tx(q_full, AF_QTILE);
        tma_load_2d(smem + AF2_SMEM_Q, &tmQ, q_full, h * HD, b * a.N + q0);
        for (int j = 0; j < NS && j < nkv; ++j) {
            mbar_expect_tx(&k_full[j], AF_KVTILE);
            tma_load_2d(smem + AF2_SMEM_K + j * AF_KVTILE, &tmKV, &k_full[j], (a.H + h) * HD, b * a.N + j * AF_BKV);
            mbar_expect_tx(&v_full[j], AF_KVTILE);
            tma_load_2d(smem + AF2_SMEM_V + j * AF_KVTILE, &tmKV, &v_full[j], (2 * a.H + h) * HD, b * a.N + j * AF_BKV);
        }
    }
    if (warp == 9) tmem_alloc<AF2_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 64;   // O_g at tmem_o + 64 g

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int j = NS; j < nkv; ++j) {   // the first NS tiles were issued in the prologue
                const int s = j % NS;
                mbar_wait_backoff(&kv_empty[s], ((j / NS) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], AF_KVTILE);
                tma_load_2d(smem + AF2_SMEM_K + s * AF_KVTILE, &tmKV, &k_full[s], (a.H + h) * HD, b * a.N + j * AF_BKV);
                mbar_expect_tx(&v_full[s], AF_KVTILE);
                tma_load_2d(smem + AF2_SMEM_V + s * AF_KVTILE, &tmKV, &v_full[s], (2 * a.H + h) * HD,
                            b * a.N + j * AF_BKV);
            }
        }
        // L2 prefetch for the CTA that will run one wave later on this SM slot: its first TMA loads then hit L2 instead
        // of paying the cold HBM round trip (the prologue is ~25 % of a CTA's life on a 197-token image)
        if (lane == 1 && a.pf_dist > 0) {
            const long long lin = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z) +
                                  a.pf_dist;
            if (lin < (long long)gridDim.x * gridDim.y * gridDim.z) {
                const int pq = (int)(lin % gridDim.x), phb = (int)(lin / gridDim.x);
                const int ph = phb % (int)gridDim.y, pb = phb / (int)gridDim.y;
                tma_prefetch_l2_2d(&tmQ, ph * HD, pb * a.N + pq * AF_BQ);
                if (pq == 0) {   // one CTA per (head, image) fetches K/V
                    for (int j = 0; j < nkv; ++j) {
                        tma_prefetch_l2_2d(&tmKV, (a.H + ph) * HD, pb * a.N + j * AF_BKV);
                        tma_prefetch_l2_2d(&tmKV, (2 * a.H + ph) * HD, pb * a.N + j * AF_BKV);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + AF2_SMEM_Q), p_addr = smem_u32(smem + AF2_SMEM_P);
            // VITK_ATTN_MMA_SPIN=1: poll without back-off (A/B runs and traces)
            const bool spin = a.mma_spin != 0;
            auto wait = [&](uint64_t* bar, uint32_t parity) {
                if (spin) mbar_wait(bar, parity);
                else mbar_wait_backoff(bar, parity);
            };
            auto issue_s = [&](int j) {
                const int s = j % NS;
                const int valid = min(AF_BKV, a.N - j * AF_BKV);
                const int ncols = (valid + 15) & ~15;
                wait(&k_full[s], (j / NS) & 1);
                if (j > 0) wait(s_free, (j - 1) & 1);
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_bf16(128, ncols, 0, 0);
                const uint64_t adesc = make_smem_desc_sw128(q_addr, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + AF2_SMEM_K + s * AF_KVTILE), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                umma_commit(s_full);
            };
            VITK_TRACE_EV(a.trace, 1);
            wait(q_full, 0);
            issue_s(0);
            VITK_TRACE_EV(a.trace, 2);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1, ks = j % NS;
                if (j + 1 < nkv) issue_s(j + 1);
                const int valid = min(AF_BKV, a.N - j * AF_BKV);
                const int ksteps = (valid + 15) >> 4;
                wait(&v_full[ks], (j / NS) & 1);
                // O_g[128, HD] += P_j[128, 32g..32g+31] * V_j[32g..32g+31, HD]: A = P (K-major), B = V read MN-major
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t v_addr = smem_u32(smem + AF2_SMEM_V + ks * AF_KVTILE);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    wait(&p_full[2 * g + s], (j >> 1) & 1);
                    if (j < 4) VITK_TRACE_EV(a.trace, (g == 0 ? 56 : 3) + j);
                    tc_fence_after_sync();
                    for (int k = 2 * g; k < ksteps && k < 2 * g + 2; ++k) {
                        const uint64_t adesc = make_smem_desc_sw128(p_addr + s * AF_QTILE + k * 32, 0, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(v_addr + k * 2048, AF_BKV * 128, 1024);
                        umma_bf16(tmem_o + 64 * g, adesc, bdesc, idesc_pv, (j > 0 || k > 2 * g) ? 1u : 0u);
                    }
                }
                umma_commit(&o_full[s]);
                umma_commit(&kv_empty[ks]);
                if (j < 8) VITK_TRACE_EV(a.trace, 8 + j);
            }
        }
    } else {
        // ===================== softmax: group g, thread == query row =====================
        const int g = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const uint32_t tmem_og = tmem_o + 64 * g;
        const bool warp_live = (q0 + wq * 32) < a.N;   // warp-uniform: at least one real query row in this warp
        float m_ref = -INFINITY, l_run = 0.f;
        const int sw = row & 7;

        for (int j = 0; j < nkv; ++j) {
            const int valid = min(AF_BKV, a.N - j * AF_BKV);
            const int nv = min(32, valid - 32 * g);       // valid columns of this group in tile j (may be <= 0)
            const bool work = warp_live && nv > 0;        // warp-uniform
            mbar_wait(s_full, j & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 0 && j < 8) VITK_TRACE_EV(a.trace, 24 + 2 * j);
            if (threadIdx.x == 128 && j < 4) VITK_TRACE_EV(a.trace, 16 + 2 * j);
            uint32_t sr[32];
            if (work) {
                tmem_ld_32x32b_x32(tmem_s + lane_off + 32 * g, sr);
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            mbar_arrive(s_free);
            if (threadIdx.x == 0 && j < 2) VITK_TRACE_EV(a.trace, 40 + 4 * j);
            if (work) {
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (nv == 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sr[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        mx4[i & 3] = fmaxf(mx4[i & 3], (i < nv) ? __uint_as_float(sr[i]) : -INFINITY);
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                const float mxs = mx * a.scale_log2;
                if (j == 0) {
                    m_ref = mxs;    // (a group with no valid column in tile 0 has none in any tile)
                } else if (__any_sync(0xffffffffu, mxs > m_ref + AF_RESCALE_TAU)) {
                    // lazy rescale (warp-uniform branch): raise the reference maximum, rescale l and O_g in TMEM
                    mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // PV_{j-1} has landed
                    tc_fence_after_sync();
                    const float m_new = fmaxf(m_ref, mxs);
                    const float alpha = ex2_approx(m_ref - m_new);
                    l_run *= alpha;
                    m_ref = m_new;
#pragma unroll
                    for (int c = 0; c < HD / 16; ++c) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tmem_og + lane_off + c * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x16(tmem_og + lane_off + c * 16, r);
                    }
                    tmem_st_wait();
                    tc_fence_before_sync();
                }
                // P buffer (j & 1) was last read by PV_{j-2}
                if (j >= 2) mbar_wait(&o_full[j & 1], ((j - 2) >> 1) & 1);
                if (threadIdx.x == 0 && j < 2) VITK_TRACE_EV(a.trace, 41 + 4 * j);
                uint8_t* p_row = smem + AF2_SMEM_P + (j & 1) * AF_QTILE + row * 128;
                float ps4[4] = {0.f, 0.f, 0.f, 0.f};
                const float nm = -m_ref;
                if (nv == 32) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            p[i] = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, nm));
                            ps4[i & 3] += p[i];
                        }
                        *reinterpret_cast<uint4*>(p_row + (((4 * g + u) ^ sw) << 4)) = make_uint4(
                            pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u < ((nv + 15) >> 4) * 2) {   // 16-byte units of 8 columns that the PV MMA may read
                            float p[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float v = ex2_approx(fmaf(__uint_as_float(sr[u * 8 + i]), a.scale_log2, nm));
                                p[i] = (u * 8 + i < nv) ? v : 0.f;
                                ps4[i & 3] += p[i];
                            }
                            *reinterpret_cast<uint4*>(p_row + (((4 * g + u) ^ sw) << 4)) = make_uint4(
                                pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                pack_bf16(p[6], p[7]));
                        }
                    }
                }
                l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
                if (threadIdx.x == 0 && j < 2) VITK_TRACE_EV(a.trace, 42 + 4 * j);
                fence_proxy_async_smem();
                if (threadIdx.x == 0 && j < 2) VITK_TRACE_EV(a.trace, 43 + 4 * j);
            }
            mbar_arrive(&p_full[2 * g + (j & 1)]);
            if (threadIdx.x == 0 && j < 8) VITK_TRACE_EV(a.trace, 25 + 2 * j);
            if (threadIdx.x == 128 && j < 4) VITK_TRACE_EV(a.trace, 17 + 2 * j);
            if (lane == 0 && j == 1) VITK_TRACE_EV(a.trace, 48 + warp);      // every softmax warp: end of tile 1
        }
        // merge the two groups' partial softmax states, then each group writes half of the head's columns
        // (the Q tile is dead: its last reader, S of the last tile, completed before this group's last s_full wait)
        float2* stat = reinterpret_cast<float2*>(smem + AF2_SMEM_Q);
        stat[g * 128 + row] = make_float2(m_ref, l_run);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp_live) {
            const float2 other = stat[(g ^ 1) * 128 + row];
            const float m0 = g ? other.x : m_ref, l0 = g ? other.y : l_run;
            const float m1 = g ? m_ref : other.x, l1 = g ? l_run : other.y;
            const float m = fmaxf(m0, m1);
            const bool has1 = a.N > 32;                     // group 1 saw at least one key: O_1 holds data
            const float w0 = ex2_approx(m0 - m);
            const float w1 = has1 ? ex2_approx(m1 - m) : 0.f;
            const float l = l0 * w0 + l1 * w1;
            const float inv = 1.0f / l;
            const float c0 = w0 * inv, c1 = w1 * inv;
            mbar_wait(&o_full[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 60);
            const int n = q0 + row;
            __nv_bfloat16* dst = a.out + ((long long)b * a.N + n) * a.D + h * HD;
            // d = 64: the tile leaves through a swizzled staging tile (P buffer 0, free once the last PV has completed)
            // and ONE TMA store; d = 48: direct 16-byte stores
            uint8_t* stage_row = smem + AF2_SMEM_P + row * 128;
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                if ((c & 1) != g) continue;
                uint32_t r0[16], r1[16];
                tmem_ld_32x32b_x16(tmem_o + lane_off + c * 16, r0);             // warp-collective
                if (has1) tmem_ld_32x32b_x16(tmem_o + 64 + lane_off + c * 16, r1);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    f[i] = has1 ? fmaf(__uint_as_float(r1[i]), c1, __uint_as_float(r0[i]) * c0)
                                : __uint_as_float(r0[i]) * c0;
                const uint4 lo = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                            pack_bf16(f[6], f[7]));
                const uint4 hi = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                            pack_bf16(f[14], f[15]));
                if constexpr (HD == 64) {
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c) ^ sw) << 4)) = lo;
                    *reinterpret_cast<uint4*>(stage_row + (((2 * c + 1) ^ sw) << 4)) = hi;
                } else if (n < a.N) {
                    st_v4(dst + c * 16, lo);
                    st_v4(dst + c * 16 + 8, hi);
                }
            }
            if (g == 0 && n < a.N) a.lse2[((long long)b * a.H + h) * a.N + n] = m + log2f(l);
            if constexpr (HD == 64) fence_proxy_async_smem();
        }
        if constexpr (HD == 64) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 0) {
                tma_store_3d(&tmO, smem + AF2_SMEM_P, h * HD, q0, b);   // rows >= N are clipped by the map
                tma_store_commit();
                tma_store_wait_read<0>();
            }
        }
        if (threadIdx.x == 0) VITK_TRACE_EV(a.trace, 61);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<AF2_TMEM_COLS>(tmem_base);
    }
}

// 2-D map over a token-major bf16 tensor viewed as [B*N rows, G*d columns] (G = 3H for qkv, H for O / dO), box
// {64 columns, box_rows}. Rank-2 boxes are markedly cheaper for the TMA unit than the rank-4 {d, G, N, B} form; the
// price is that a tile may run past its image into the next one (rows) or the next head (columns, d = 48): every
// consumer masks those rows / columns (softmax columns >= N, k-steps and UMMA N limited to d), and rows past the end
// of the tensor are zero-filled by TMA.
int make_tok_tmap2d(CUtensorMap* out, const void* p, long long rows, long long cols, int box_rows) {
    return make_tmap_2d_bf16(out, p, (uint64_t)cols, (uint64_t)rows, (uint64_t)cols, 64, (uint32_t)box_rows);
}

// VITK_ATTN_WG2=0 selects the one-group kernels (A/B runs); read on every call so a process can time both.
// VITK_ATTN_WG2=0 selects the one-group kernels (A/B runs); read on every call so a process can time both.
bool attn_two_groups() {
    const char* e = getenv("VITK_ATTN_WG2");
    return !(e && e[0] == '0');
}

// CTA i prefetches for CTA i + 2 x #SM (the next occupant of its slot); VITK_ATTN_PREFETCH=0 turns it off (A/B runs)
int attn_prefetch_dist() {
    const char* e = getenv("VITK_ATTN_PREFETCH");
    return (e && e[0] == '0') ? 0 : 2 * sm_count();
}

template <int HD>
static int launch_attn_fwd(const void* qkv, const AttnFwdArgs& a, int B, int N, int H, int d, cudaStream_t st) {
    CUtensorMap tmq, tmkv;
    if (make_tok_tmap2d(&tmq, qkv, (long long)B * N, 3LL * H * d, 128) ||
        make_tok_tmap2d(&tmkv, qkv, (long long)B * N, 3LL * H * d, AF_BKV))
        return VITK_ERR_TMAP;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES) !=
            cudaSuccess)
            return VITK_ERR_CUDA;
        attr = true;
    }
    dim3 grid((N + AF_BQ - 1) / AF_BQ, H, B);
    if (attn_two_groups()) {
        CUtensorMap tmo;
        if (make_tmap_3d_tok_store(&tmo, a.out, (uint64_t)H * d, (uint64_t)N, (uint64_t)B, 128)) return VITK_ERR_TMAP;
        static bool attr2 = false;
        if (!attr2) {
            if (cudaFuncSetAttribute(attn_fwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AF2_SMEM_BYTES) != cudaSuccess)
                return VITK_ERR_CUDA;
            attr2 = true;
        }
        attn_fwd2_kernel<HD><<<grid, AF2_THREADS, AF2_SMEM_BYTES, st>>>(tmq, tmkv, tmo, a);
    } else {
        attn_fwd_kernel<HD><<<grid, AF_THREADS, AF_SMEM_BYTES, st>>>(tmq, tmkv, a);
    }
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_attn_fwd(const void* qkv_bf16, void* out_bf16, float* lse2, int B, int N, int H, int d,
                             float scale, void* stream) {
    if (B <= 0 || N <= 0 || H <= 0 || !(d == 64 || d == 48) || !qkv_bf16 || !out_bf16 || !lse2) return VITK_ERR_ARG;
    AttnFwdArgs a;
    a.B = B; a.H = H; a.N = N; a.D = H * d;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    a.lse2 = lse2;
    a.trace = nullptr;
    a.pf_dist = attn_prefetch_dist();
    {
        static int spin = -1;
        if (spin < 0) { const char* e = getenv("VITK_ATTN_MMA_SPIN"); spin = (e && e[0] == '1') ? 1 : 0; }
        a.mma_spin = spin;
    }
#ifdef VITK_TRACE
    a.trace = g_attn_trace;
#endif
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d == 64) return launch_attn_fwd<64>(qkv_bf16, a, B, N, H, d, st);
    return launch_attn_fwd<48>(qkv_bf16, a, B, N, H, d, st);
}

#ifdef VITK_TRACE
// instrumented build only: device buffer of int64 [n_ctas * 64] that the attention kernels fill with SM clocks
extern "C" int vitk_debug_set_trace(void* p) {
    g_attn_trace = reinterpret_cast<long long*>(p);
    return VITK_OK;
}
#endif
