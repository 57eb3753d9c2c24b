// Short-sequence per-(image, head) products of the talking-heads attention path (models/cait.py:111-128, and the
// autograd of the same), N <= 208 tokens (CaiT at 224 px: 196), head dim 48 or 64, on tcgen05:
//
//   th_scores : S[b,h,i,j]   = sum_e A[b,i,h,e] * Bm[b,j,h,e]      (q k^T  and  dO v^T : bf16 / fp32 planes [B,H,N,Np])
//   th_apply  : O[b,i,(h,e)] = sum_j P[b,h,i,j] * X[b,j,(h,e)]     (P' v   and  dS k   : token-major bf16)
//   th_apply_t: O[b,j,(h,e)] = sum_i P[b,h,i,j] * X[b,i,(h,e)]     (P'^T dO and dS^T q : token-major bf16)
//
// The talking-heads mixes couple all heads of a (query, key) position, so the [B,H,N,N] planes exist in HBM between
// these products and the mixing kernels (th_mix2.cuh); what this file fixes is the cost of the products themselves.
// The generic batched GEMM ran them as 128 x 128 tiles with one 64-deep k-block each behind rank-4 tensor maps -- a
// latency-bound pipeline of thousands of tiny tiles (59 us per product at CaiT-S24 bs128). Here
//   * whole-image operand tiles (<= 208 rows) arrive through rank-3 maps {columns, N, image or plane} whose out-of-range
//     rows are ZERO-filled, so the contraction tail needs no masking and a row block never reads the next image;
//   * the kernels are PERSISTENT: one thread block per SM walks work items (row block, image, head or head group) with
//     the operand ring and two TMEM accumulator buffers running across item boundaries -- the first (one block per
//     item) versions sat at 24-34 % of HBM bandwidth with the SMs idle 30 % of the time (1.73 waves; ncu in
//     profiles/r02_summary.md);
//   * th_scores leaves through swizzled staging tiles and TMA stores (per-lane stores of a row-per-lane TMEM slab touch
//     32 different 128-byte lines per instruction: 50 us for an 80 MB plane), th_apply / th_apply_t write whole
//     32-byte sectors of the token rows.
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int TG_THREADS = 320;                 // warps 0-7 epilogue, warp 8 TMA, warp 9 MMA
constexpr int TG_ROWS = 208;                    // rows of a whole-image operand tile (tokens, padded to 16)
constexpr int TG_T128 = 128 * 128;              // [128 rows x 64 bf16] swizzled tile bytes
constexpr int TG_TIMG = TG_ROWS * 128;          // [208 rows x 64 bf16] swizzled tile bytes (26 x 1024)

struct ThGemmArgs {
    int B, H, N, Np;        // Np: row pitch of a plane (multiple of 8)
    int HG;                 // heads per work item (th_apply*)
    int a_col0, b_col0;     // first column of head 0 inside the token-major operands
    void* out;              // th_scores: plane [B,H,N,Np] (bf16 or fp32); th_apply*: token-major bf16 [B*N, ldo]
    long long ldo;
    int o_col0;
    int out_f32;
    float* colsum;          // th_apply*: optional fp32 [H*d], += column sums of the output (bias gradient of the qkv Linear)
};

__device__ __forceinline__ uint4 pack8_bf16(const uint32_t* v) {
    return make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])),
                      pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                      pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])),
                      pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
}

// ------------------------------------------------------------------------------------------------------------------
// th_scores: work item = (row block, image, head), row block slowest so that every thread block gets its share of the
// short last row block. Stage = A tile [128 x 64] + B tile [208 x 64]; two accumulators [128 x 208] fp32 in TMEM.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TS_NS = 3;
constexpr int TS_STAGE = TG_T128 + TG_TIMG;                 // 43008
constexpr int TS_SMEM_OUT = TS_NS * TS_STAGE;               // 2 x [128 x 64] bf16 staging tiles of the TMA stores
constexpr int TS_SMEM_BAR = TS_SMEM_OUT + 2 * TG_T128;
constexpr int TS_SMEM_BYTES = TS_SMEM_BAR + 128;

template <int HD>
__global__ void __launch_bounds__(TG_THREADS, 1)
th_scores_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const ThGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TS_SMEM_BAR);
    uint64_t* full = bars;                   // [TS_NS]
    uint64_t* empty = bars + TS_NS;          // [TS_NS]
    uint64_t* acc_full = bars + 2 * TS_NS;   // [2]
    uint64_t* acc_free = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk16 = (a.N + 15) & ~15;                     // UMMA N (<= 208)
    const int per_rb = a.B * a.H;
    const int items = ((a.N + 127) >> 7) * per_rb;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (!a.out_f32) tma_prefetch_desc(&tmO);
        for (int i = 0; i < TS_NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 256);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int rb = w / per_rb, bh = w - rb * per_rb;
                const int b = bh / a.H, h = bh - b * a.H;
                const int s = it % TS_NS;
                if (it >= TS_NS) mbar_wait_backoff(&empty[s], ((it / TS_NS) & 1) ^ 1);
                uint8_t* st = smem + s * TS_STAGE;
                mbar_expect_tx(&full[s], TS_STAGE);
                tma_load_3d(st, &tmA, &full[s], a.a_col0 + h * HD, rb * 128, b);
                tma_load_3d(st + TG_T128, &tmB, &full[s], a.b_col0 + h * HD, 0, b);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)nk16, 0, 0);
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int s = it % TS_NS, as = it & 1;
                mbar_wait_backoff(&full[s], (it / TS_NS) & 1);
                if (it >= 2) mbar_wait_backoff(&acc_free[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + s * TS_STAGE);
                const uint64_t adesc = make_smem_desc_sw128(sa, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(sa + TG_T128, 0, 1024);
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                umma_commit(&empty[s]);
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        // epilogue: lane quadrant = rows; warp half = columns
        const int q = warp & 3, half = warp >> 2;
        const uint32_t lane_off = uint32_t(q * 32) << 16;
        const int r = q * 32 + lane;                           // row inside the tile
        int it = 0;
        int nstore = 0;                                        // TMA stores issued so far (staging tile = nstore & 1)
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int rb = w / per_rb, bh = w - rb * per_rb;
            const int as = it & 1;
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * 256) + lane_off;
            mbar_wait(&acc_full[as], (it >> 1) & 1);
            tc_fence_after_sync();
            if (!a.out_f32) {
                // bf16 planes leave as [128 rows x 64 columns] tiles: the eight warps transpose their TMEM slabs into a
                // swizzled staging tile and ONE TMA store writes it (rows >= N and columns >= Np are clipped by the map)
                const int sw = r & 7;
                const int nchunk = (min(nk16, a.Np) + 63) >> 6;
                for (int c = 0; c < nchunk; ++c, ++nstore) {
                    const int col0 = 64 * c + 32 * half;
                    const bool active = col0 < nk16;           // (warp-uniform)
                    uint32_t v[32];
                    if (active) {
                        tmem_ld_32x32b_x32(tmem_acc + col0, v);
                        tmem_ld_wait();
                    }
                    if (c == nchunk - 1) {                     // all TMEM reads of this accumulator are done
                        tc_fence_before_sync();
                        mbar_arrive(&acc_free[as]);
                    }
                    uint8_t* stg = smem + TS_SMEM_OUT + (nstore & 1) * TG_T128;
                    // the store issued two steps ago (out of the same staging tile) has finished reading it
                    if (threadIdx.x == 0) tma_store_wait_read<1>();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    uint8_t* stg_row = stg + r * 128;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint4 pk = active ? pack8_bf16(v + 8 * u) : make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(stg_row + (((4 * half + u) ^ sw) << 4)) = pk;
                    }
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 0) {
                        tma_store_3d(&tmO, stg, 64 * c, rb * 128, bh);
                        tma_store_commit();
                    }
                }
            } else {
                // fp32 planes (VITK_TH_S_F32=1): direct 16-byte stores, columns [0,112) / [112,208) per warp half
                const int row = rb * 128 + r;
                const bool row_ok = row < a.N;
                const int c_begin = half ? 112 : 0, c_end = half ? TG_ROWS : 112;
                float* orow = reinterpret_cast<float*>(a.out) + ((long long)bh * a.N + row) * (long long)a.Np;
                for (int c = c_begin; c < c_end && c < nk16; c += 16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_acc + c, v);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + 4 * u + 4 <= a.Np)
                                st_v4(orow + c + 4 * u, make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&acc_free[as]);
            }
        }
        if (threadIdx.x == 0) tma_store_wait<0>();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// th_apply (TRANS = false): work item = (block of 128 rows i, image, head group). Stage (one head) = P rows
//   [128 x 256 columns] as four K-major tiles + X_h [208 x 64] (MN-major B operand).
// th_apply_t (TRANS = true): work item = (block of 128 columns j, image, head group). Stage = P columns [208 rows x 128
//   columns] as two MN-major tiles + X_h [208 x 64] (MN-major B operand).
// The outputs of an item's heads sit side by side in one of two TMEM accumulator buffers (HG * HD <= 256 columns) and
// leave as 32-byte pieces of the token rows while the next item's operands stream in.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TA_NS = 2;
template <bool TRANS> struct TaCfg {
    static constexpr int A_BYTES = TRANS ? 2 * TG_TIMG : 4 * TG_T128;      // 53248 / 65536
    static constexpr int STAGE = A_BYTES + TG_TIMG;
    static constexpr int SMEM_BAR = TA_NS * STAGE;
    static constexpr int SMEM_BYTES = SMEM_BAR + 128;
};

template <int HD, bool TRANS>
__global__ void __launch_bounds__(TG_THREADS, 1)
th_apply_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmX, const ThGemmArgs a) {
    using Cfg = TaCfg<TRANS>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + TA_NS;
    uint64_t* acc_full = bars + 2 * TA_NS;   // [2]
    uint64_t* acc_free = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nks = (a.N + 15) >> 4;                        // k-steps over the contracted token axis (<= 13)
    const int nhg = (a.H + a.HG - 1) / a.HG;
    const int per_mb = a.B * nhg;
    const int items = ((a.N + 127) >> 7) * per_mb;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < TA_NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 256);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            int it = 0;                                     // heads loaded so far (position in the stage ring)
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                const int mb = w / per_mb, rem = w - mb * per_mb;
                const int b = rem / nhg, hg = rem - b * nhg;
                const int m0 = mb * 128;
                const int h0 = hg * a.HG, h1 = min(a.H, h0 + a.HG);
                for (int h = h0; h < h1; ++h, ++it) {
                    const int s = it % TA_NS;
                    if (it >= TA_NS) mbar_wait_backoff(&empty[s], ((it / TA_NS) & 1) ^ 1);
                    uint8_t* st = smem + s * Cfg::STAGE;
                    const int plane = b * a.H + h;
                    if constexpr (!TRANS) {
                        // columns past Np and rows past N arrive as zeros; tiles that start past Np are not fetched
                        const int nchunk = min(4, (a.Np + 63) >> 6);
                        mbar_expect_tx(&full[s], nchunk * TG_T128 + TG_TIMG);
                        for (int c = 0; c < nchunk; ++c) tma_load_3d(st + c * TG_T128, &tmP, &full[s], 64 * c, m0, plane);
                    } else {
                        const int nchunk = (m0 + 64 < a.Np) ? 2 : 1;
                        mbar_expect_tx(&full[s], nchunk * TG_TIMG + TG_TIMG);
                        for (int c = 0; c < nchunk; ++c)
                            tma_load_3d(st + c * TG_TIMG, &tmP, &full[s], m0 + 64 * c, 0, plane);
                    }
                    tma_load_3d(st + Cfg::A_BYTES, &tmX, &full[s], a.b_col0 + h * HD, 0, b);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, HD, TRANS ? 1u : 0u, 1u);
            int it = 0, ai = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++ai) {
                const int hg = (w % per_mb) % nhg;
                const int h0 = hg * a.HG, h1 = min(a.H, h0 + a.HG);
                const int as = ai & 1;
                if (ai >= 2) mbar_wait_backoff(&acc_free[as], ((ai >> 1) & 1) ^ 1);
                for (int h = h0; h < h1; ++h, ++it) {
                    const int s = it % TA_NS;
                    mbar_wait_backoff(&full[s], (it / TA_NS) & 1);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * Cfg::STAGE), sb = sa + Cfg::A_BYTES;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256 + (h - h0) * HD);
                    for (int k = 0; k < nks; ++k) {
                        uint64_t adesc;
                        if constexpr (!TRANS)   // K-major: k-step k = 16 columns = tile k / 4, +32 B per step inside the tile
                            adesc = make_smem_desc_sw128(sa + (k >> 2) * TG_T128 + (k & 3) * 32, 0, 1024);
                        else                    // MN-major: two 64-wide column chunks TG_TIMG apart, +16 rows per k-step
                            adesc = make_smem_desc_sw128(sa + k * 2048, TG_TIMG, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(sb + k * 2048, TG_TIMG, 1024);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        const int q = warp & 3, half = warp >> 2;
        const uint32_t lane_off = uint32_t(q * 32) << 16;
        int ai = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++ai) {
            const int mb = w / per_mb, rem = w - mb * per_mb;
            const int b = rem / nhg, hg = rem - b * nhg;
            const int h0 = hg * a.HG, h1 = min(a.H, h0 + a.HG);
            const int as = ai & 1;
            const int row = mb * 128 + q * 32 + lane;
            const bool row_ok = row < a.N;
            const int nch = ((h1 - h0) * HD) >> 4;              // 16-column chunks of this item's output
            const int ch0 = half ? (nch + 1) / 2 : 0, ch1 = half ? nch : (nch + 1) / 2;
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(a.out) + ((long long)b * a.N + row) * a.ldo +
                                  a.o_col0 + h0 * HD;
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * 256) + lane_off;
            mbar_wait(&acc_full[as], (ai >> 1) & 1);
            tc_fence_after_sync();
            auto store16 = [&](int ch, const uint32_t* v) {      // 16 columns of this lane's row -> 32 bytes of bf16
                if (row_ok) {
                    st_v4(orow + ch * 16, pack8_bf16(v));
                    st_v4(orow + ch * 16 + 8, pack8_bf16(v + 8));
                }
                if (a.colsum != nullptr) {                       // (warp-uniform) 16-shuffle reduction over the 32 rows
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = row_ok ? __uint_as_float(v[i]) : 0.f;
                    const float tot = warp_colsum16(f, lane);
                    if ((lane & 1) == 0) atomicAdd(a.colsum + h0 * HD + ch * 16 + warp_colsum16_col(lane), tot);
                }
            };
            int ch = ch0;
            for (; ch + 2 <= ch1; ch += 2) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + ch * 16, v);
                tmem_ld_wait();
                store16(ch, v);
                store16(ch + 1, v + 16);
            }
            if (ch < ch1) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_acc + ch * 16, v);
                tmem_ld_wait();
                store16(ch, v);
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_free[as]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// rank-3 bf16 map {columns, rows, slabs} with byte strides {row pitch, slab pitch}; box {64 columns, box_rows, 1}.
static int make_tmap_3d_rows(CUtensorMap* out, const void* p, uint64_t cols, uint64_t rows, uint64_t slabs,
                             uint64_t row_pitch_elems, uint64_t slab_pitch_elems, uint32_t box_rows) {
    uint64_t dims[3] = {cols, rows, slabs};
    uint64_t strides[2] = {row_pitch_elems * 2, slab_pitch_elems * 2};
    uint32_t box[3] = {64, box_rows, 1};
    return make_tmap(out, p, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

static bool th_gemm_shape_ok(int B, int N, int H, int d, int Np) {
    return B > 0 && H > 0 && N > 0 && N <= TG_ROWS && (d == 48 || d == 64) && Np >= N && (Np % 8) == 0 && Np <= 256;
}

template <int HD>
static int launch_th_scores(const void* A, long long lda, int a_cols, const void* Bm, long long ldb, int b_cols,
                            const ThGemmArgs& a, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (make_tmap_3d_rows(&tmA, A, (uint64_t)a_cols, (uint64_t)a.N, (uint64_t)a.B, (uint64_t)lda, (uint64_t)lda * a.N, 128) ||
        make_tmap_3d_rows(&tmB, Bm, (uint64_t)b_cols, (uint64_t)a.N, (uint64_t)a.B, (uint64_t)ldb, (uint64_t)ldb * a.N, TG_ROWS))
        return VITK_ERR_TMAP;
    CUtensorMap tmO;
    memset(&tmO, 0, sizeof(tmO));
    if (!a.out_f32 && make_tmap_3d_rows(&tmO, a.out, (uint64_t)a.Np, (uint64_t)a.N, (uint64_t)a.B * a.H, (uint64_t)a.Np,
                                        (uint64_t)a.Np * a.N, 128))
        return VITK_ERR_TMAP;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(th_scores_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES) !=
            cudaSuccess)
            return VITK_ERR_CUDA;
        attr = true;
    }
    const long long items = (long long)((a.N + 127) / 128) * a.B * a.H;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    th_scores_kernel<HD><<<grid, TG_THREADS, TS_SMEM_BYTES, st>>>(tmA, tmB, tmO, a);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

template <int HD, bool TRANS>
static int launch_th_apply(const void* P, const void* X, long long ldx, int x_cols, const ThGemmArgs& a, cudaStream_t st) {
    CUtensorMap tmP, tmX;
    if (make_tmap_3d_rows(&tmP, P, (uint64_t)a.Np, (uint64_t)a.N, (uint64_t)a.B * a.H, (uint64_t)a.Np,
                          (uint64_t)a.Np * a.N, TRANS ? TG_ROWS : 128) ||
        make_tmap_3d_rows(&tmX, X, (uint64_t)x_cols, (uint64_t)a.N, (uint64_t)a.B, (uint64_t)ldx, (uint64_t)ldx * a.N, TG_ROWS))
        return VITK_ERR_TMAP;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(th_apply_kernel<HD, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TaCfg<TRANS>::SMEM_BYTES) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr = true;
    }
    const long long items = (long long)((a.N + 127) / 128) * a.B * ((a.H + a.HG - 1) / a.HG);
    const int grid = (int)(items < sm_count() ? items : sm_count());
    th_apply_kernel<HD, TRANS><<<grid, TG_THREADS, TaCfg<TRANS>::SMEM_BYTES, st>>>(tmP, tmX, a);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

// Heads per work item of th_apply*: HG * d <= 256 TMEM columns, and the item count should split evenly over the SMs
// (a block that gets 4 items while its neighbours get 3 sets the kernel's time). More heads per item are preferred
// unless fewer heads balance at least 3 % better. VITK_TH_APPLY_HG overrides (A/B runs).
static int th_apply_heads_per_item(int B, int N, int H, int d) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("VITK_TH_APPLY_HG"); forced = e ? atoi(e) : 0; }
    const int hmax = min(H, 256 / d);
    if (forced > 0) return min(forced, hmax);
    int best = hmax;
    double best_eff = -1.0;
    for (int hg = hmax; hg >= 1; --hg) {
        const long long items = (long long)((N + 127) / 128) * B * ((H + hg - 1) / hg);
        const long long rounds = (items + sm_count() - 1) / sm_count();
        const double eff = (double)items / (double)(rounds * sm_count());
        if (eff > best_eff + 0.03) { best_eff = eff; best = hg; }
    }
    return best;
}

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_th_gemm_supported(int N, int d, int Np) {
    return th_gemm_shape_ok(1, N, 1, d, Np) ? 1 : 0;
}

extern "C" int vitk_th_scores(const void* a_bf16, long long lda, int a_cols, int a_col0, const void* b_bf16,
                              long long ldb, int b_cols, int b_col0, void* out, int out_f32, int B, int N, int H, int d,
                              int Np, void* stream) {
    if (!a_bf16 || !b_bf16 || !out || !th_gemm_shape_ok(B, N, H, d, Np)) return VITK_ERR_ARG;
    if ((lda % 8) != 0 || (ldb % 8) != 0 || a_col0 < 0 || b_col0 < 0 || a_col0 + H * d > a_cols || b_col0 + H * d > b_cols ||
        a_cols > lda || b_cols > ldb)
        return VITK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(a_bf16) | reinterpret_cast<uintptr_t>(b_bf16) | reinterpret_cast<uintptr_t>(out)) & 15)
        return VITK_ERR_ARG;
    ThGemmArgs a;
    a.B = B; a.H = H; a.N = N; a.Np = Np;
    a.HG = 1;
    a.a_col0 = a_col0; a.b_col0 = b_col0;
    a.out = out; a.ldo = Np; a.o_col0 = 0; a.out_f32 = out_f32 ? 1 : 0;
    a.colsum = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d == 64) return launch_th_scores<64>(a_bf16, lda, a_cols, b_bf16, ldb, b_cols, a, st);
    return launch_th_scores<48>(a_bf16, lda, a_cols, b_bf16, ldb, b_cols, a, st);
}

extern "C" int vitk_th_apply(const void* p_bf16, const void* x_bf16, long long ldx, int x_cols, int x_col0, void* out_bf16,
                             long long ldo, int o_col0, int transpose, float* colsum, int B, int N, int H, int d, int Np,
                             void* stream) {
    if (!p_bf16 || !x_bf16 || !out_bf16 || !th_gemm_shape_ok(B, N, H, d, Np)) return VITK_ERR_ARG;
    if ((ldx % 8) != 0 || (ldo % 8) != 0 || (o_col0 % 8) != 0 || x_col0 < 0 || x_col0 + H * d > x_cols || x_cols > ldx ||
        o_col0 < 0 || o_col0 + H * d > ldo)
        return VITK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(p_bf16) | reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15)
        return VITK_ERR_ARG;
    ThGemmArgs a;
    a.B = B; a.H = H; a.N = N; a.Np = Np;
    a.HG = th_apply_heads_per_item(B, N, H, d);
    a.a_col0 = 0; a.b_col0 = x_col0;
    a.out = out_bf16; a.ldo = ldo; a.o_col0 = o_col0; a.out_f32 = 0;
    a.colsum = colsum;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d == 64) {
        if (transpose) return launch_th_apply<64, true>(p_bf16, x_bf16, ldx, x_cols, a, st);
        return launch_th_apply<64, false>(p_bf16, x_bf16, ldx, x_cols, a, st);
    }
    if (transpose) return launch_th_apply<48, true>(p_bf16, x_bf16, ldx, x_cols, a, st);
    return launch_th_apply<48, false>(p_bf16, x_bf16, ldx, x_cols, a, st);
}
