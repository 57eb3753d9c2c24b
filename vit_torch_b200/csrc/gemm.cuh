// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (128B swizzle) -> shared-memory ring -> tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulators in TMEM,
//   double-buffered) -> tcgen05.ld epilogue with fused bias / GELU / residual / LayerScale / GELU-grad / split-K.
//
//   D[M,N] = A[M,K] * B[N,K]^T            (A and B independently K-major or MN-major in global memory)
//
// Replaces on the reference path: nn.Linear forward (models/cait.py:99,102,113,126; timm/DINO Attention.qkv/proj,
// Mlp.fc1/fc2 -- in-repo witness models/swin.py:24-30) and the autograd dgrad / wgrad of the same Linears.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 4..11 = epilogue (two warps per TMEM lane quarter = warp % 4, each taking half of the BN columns).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace vitk {

enum GemmEpilogue : int {
    EPI_STORE_BF16 = 0,  // out_bf16 = acc (+bias)
    EPI_BIAS_GELU = 1,   // pre = acc + bias ; out_bf16 = gelu'(pre) (saved for backward) ; out2_bf16 = gelu(pre)
    EPI_RESID_F32 = 2,   // v = acc + bias ; [out2_bf16 = v] ; out_f32 = resid + gamma * v   (gamma optional)
    EPI_DGELU = 3,       // out_bf16 = acc * aux_bf16   (aux = gelu'(pre) saved by EPI_BIAS_GELU)
    EPI_ATOMIC_F32 = 4,  // out_f32 += acc   (red.global.add; split-K wgrad accumulates into the fp32 grad)
    EPI_STORE_F32 = 5,   // out_f32 = acc (+bias)
    EPI_TOKENS_F32 = 6,  // PatchEmbed: row r = (b, p) -> out_f32[b*tok_N + tok_T + p] = acc + bias + pos[tok_T + p]
};

struct GemmArgs {
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    int splits, kblocks_per_split, num_kblocks;
    const float* bias;   // [N] fp32 or null
    const float* gamma;  // [N] fp32 or null
    const float* resid;  // fp32 [M, ldr] or null
    long long ldr;
    void* out;  // primary output
    long long ldo;
    void* out2;  // secondary bf16 output or null
    long long ldo2;
    const __nv_bfloat16* aux;  // bf16 [M, ldaux] (EPI_DGELU)
    long long ldaux;
    const float* rowscale;  // per-sample scale (DropPath mask / keep_prob) indexed by row / rows_per_sample, or null
    int rows_per_sample;
    int tok_n, tok_N, tok_T;  // EPI_TOKENS_F32: patches per image, tokens per image, prefix tokens
    // batched mode: nbatch_h * nbatch_b independent GEMMs (attention-style batches over (head, image)); operands are
    // addressed through the two outer dims of the 4-D tensor maps, outputs through element strides
    int nbatch_h, nbatch_b;
    long long so_h, so_b;    // out strides (elements) per inner / outer batch index
    int a_perm[3], b_perm[3];  // tensor-map dim 1+i takes logical coordinate perm[i] (0 = row, 1 = batch_h, 2 = batch_b)
    int tma_out;    // 1: `out` (and BIAS_GELU's second output) leave through TMA stores (tensor maps tmO / tmO2)
    int dbg;        // VITK_GEMM_DBG experiment switches (0 in production): 1 no L2 prefetch, 2 plain loads, 4 no loads, 8 no stores
    float* colsum;  // optional fp32 [N]: += column sums of the values stored to `out` (bias gradient of the next Linear)
    // EPI_TOKENS_F32 behind the im2col-free PatchEmbed kernel (patch_embed.cu): an M tile is (image, group of pe_rpt patch
    // rows), its first pe_rows rows are the patches pe_rows * group ... of that image. 0 = plain row-major patch rows.
    int pe_G, pe_rows;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_WARPS = 8;

// Outputs leave through TMA stores: each epilogue warp stages [32 rows x 64 B] (16 fp32 / 32 bf16 columns, SWIZZLE_64B)
// in its own 2 KB shared-memory buffer per output and one elected lane issues cp.async.bulk.tensor. Row-per-lane LSU
// stores cost one L1 tag cycle per 32-byte sector (32 distinct lines per warp instruction): at 1 sector/clk/SM the
// 310 MB that the fc1 epilogue writes are 37 us of LSU time, the same order as the MMA time of the tile.
constexpr int EPI_STG_BYTES = 2048;
template <int EPI> struct EpiSmem {
    static constexpr int kOutBufs = (EPI == EPI_ATOMIC_F32 || EPI == EPI_TOKENS_F32) ? 0 : (EPI == EPI_BIAS_GELU ? 2 : 1);
    static constexpr int STG_BYTES = kOutBufs * GEMM_EPI_WARPS * EPI_STG_BYTES;
};

template <int BN, int EPI> struct GemmCfg {
    static constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    static constexpr int B_STAGE_BYTES = BN * GEMM_BK * 2;       // 32 KB (BN=256) / 16 KB (BN=128)
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 256;
    static constexpr int VEC_BYTES = GEMM_EPI_WARPS * 2 * (BN / 2) * 4;  // per epilogue warp: bias | gamma of its columns
    static constexpr int STG_BYTES = EpiSmem<EPI>::STG_BYTES;
    static constexpr int FIXED_BYTES = BAR_BYTES + VEC_BYTES + STG_BYTES + 2048;  // +2 KB: two 1 KB alignments
    static constexpr int MAX_STAGES = (BN == 256) ? 4 : 6;
    static constexpr int FIT_STAGES = (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT_STAGES < MAX_STAGES ? FIT_STAGES : MAX_STAGES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
};

// ----------------------------------------------------------------------------------------------------------------
// Epilogue, executed in the TMEM-native domain: tcgen05.ld 32x32b hands lane l of a warp 16 consecutive accumulator
// columns of row (quarter * 32 + l). 16 columns are 32 B of bf16 / 64 B of fp32 per lane -- whole 32-byte sectors --
// and move with the 256-bit global loads / stores of sm_100 (LDG / STG.256): no shared-memory transpose, one row
// pointer per lane and operand, a single basic block of math per chunk (8 independent fp32x2 pairs of ILP).
// Chunks that are partial (N tail) or whose pointers are not 32-byte aligned take a 4-column path.
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg256_plain(const void* p, uint32_t* r) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// L2 prefetch of `bytes` (multiple of 16) at a 16-byte aligned global address
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int EPI> struct EpiTraits {
    static constexpr bool kBias = (EPI == EPI_STORE_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32 ||
                                   EPI == EPI_STORE_F32 || EPI == EPI_TOKENS_F32);
    static constexpr bool kGamma = (EPI == EPI_RESID_F32);
    static constexpr bool kResid = (EPI == EPI_RESID_F32 || EPI == EPI_TOKENS_F32);  // fp32 operand rows
    static constexpr bool kAux = (EPI == EPI_DGELU);                                  // bf16 operand rows
    static constexpr bool kOutF32 = (EPI == EPI_RESID_F32 || EPI == EPI_ATOMIC_F32 || EPI == EPI_STORE_F32 ||
                                     EPI == EPI_TOKENS_F32);
    static constexpr bool kColsum = (EPI == EPI_STORE_BF16 || EPI == EPI_DGELU);
    static constexpr int kOpWords = kResid ? 16 : (kAux ? 8 : 1);  // 32-bit registers of prefetched operand per chunk
    // operand chunks in flight per warp. DRAM latency is covered by the L2 prefetch of the next tile's operand rows
    // (issued a whole tile ahead); this ring covers L2 latency.
    static constexpr int kDepth = (kResid || kAux) ? 4 : 2;
};

struct EpiRow {  // per-lane row state for one tile; pointers are at (row, first column of this warp)
    bool ok;                    // row < M
    float rs;                   // DropPath row scale, 1 when absent
    void* out;                  // null when the output is skipped
    __nv_bfloat16* out2;        // null when absent
    const float* resid;         // null when absent
    const __nv_bfloat16* aux;
};

// Epilogue math on 4 consecutive columns of one row. v: accumulators; b / gm: bias / LayerScale gamma (zeros / ones when
// absent); r: fp32 operand (residual / pos-embed); ax: 4 bf16 operands (GELU'). Outputs: of (fp32 out), oh (bf16 out),
// oh2 (bf16 out2). Returns the value written to `out` (for the optional column-sum reduction).
template <int EPI, bool WANT_DGELU>
__device__ __forceinline__ float4 epi_math4(float4 v, const float4 b, const float4 gm, const float4 r, const uint2 ax,
                                            const float rs, float4& of, uint2& oh, uint2& oh2) {
    if constexpr (EpiTraits<EPI>::kBias) { v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if constexpr (EPI == EPI_STORE_BF16) {
        oh = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else if constexpr (EPI == EPI_BIAS_GELU) {
        float g0, g1, g2, g3, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        if constexpr (WANT_DGELU) {  // gelu'(pre) is only needed for backward
            gelu_pair<true>(v.x, v.y, g0, g1, d0, d1);
            gelu_pair<true>(v.z, v.w, g2, g3, d2, d3);
        } else {
            gelu_pair<false>(v.x, v.y, g0, g1, d0, d1);
            gelu_pair<false>(v.z, v.w, g2, g3, d2, d3);
        }
        oh = make_uint2(pack_bf16(d0, d1), pack_bf16(d2, d3));
        oh2 = make_uint2(pack_bf16(g0, g1), pack_bf16(g2, g3));
    } else if constexpr (EPI == EPI_RESID_F32) {
        oh2 = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        v.x *= gm.x * rs; v.y *= gm.y * rs; v.z *= gm.z * rs; v.w *= gm.w * rs;
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;  // zeros when there is no residual
        of = v;
    } else if constexpr (EPI == EPI_DGELU) {
        v.x *= bf16_lo(ax.x); v.y *= bf16_hi(ax.x); v.z *= bf16_lo(ax.y); v.w *= bf16_hi(ax.y);
        oh = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else if constexpr (EPI == EPI_TOKENS_F32) {
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        of = v;
    } else {  // EPI_ATOMIC_F32 / EPI_STORE_F32
        of = v;
    }
    return v;
}

// ----------------------------------------------------------------------------------------------------------------
// Per-warp epilogue engine shared by the 1-CTA and the 2-CTA kernels. One instance per epilogue thread; `tile()` drains
// this warp's [32 rows x BN/2 columns] slab of one accumulator stage.
// ----------------------------------------------------------------------------------------------------------------
template <int BN, int EPI> struct EpilogueWarp {
    using T = EpiTraits<EPI>;
    static constexpr int COLS_PER_WARP = BN / 2;
    static constexpr int NCHUNK = COLS_PER_WARP / 16;
    static constexpr int PF = T::kDepth < NCHUNK ? T::kDepth : NCHUNK;
    static_assert(NCHUNK % PF == 0 && PF % 2 == 0, "prefetch ring must tile the chunk loop");
    static constexpr int OUT_ESZ = T::kOutF32 ? 4 : 2;

    const GemmArgs& g;
    const int lane, quarter, half;
    float* wvec;  // per-warp staging of this tile's bias / gamma columns: [bias COLS_PER_WARP | gamma COLS_PER_WARP]
    uint32_t wvec_s;
    uint32_t stg_s;             // this warp's TMA-store staging buffer(s): [kOutBufs][32 rows x 64 B], SWIZZLE_64B
    const CUtensorMap* tm_out;  // tensor maps of `out` / BIAS_GELU's `out2` (valid when g.tma_out)
    const CUtensorMap* tm_out2;
    bool has_bias, has_gamma, has_resid, want_out, vec_ok_static;

    __device__ __forceinline__ EpilogueWarp(const GemmArgs& g_, float* svec, uint8_t* sstg, const CUtensorMap* tmo,
                                            const CUtensorMap* tmo2, int ew, int warp, int lane_)
        : g(g_), lane(lane_), quarter(warp & 3), half(ew >> 2), tm_out(tmo), tm_out2(tmo2) {
        wvec = svec + ew * (2 * COLS_PER_WARP);
        wvec_s = smem_u32(wvec);
        stg_s = smem_u32(sstg + ew * (EpiSmem<EPI>::kOutBufs * EPI_STG_BYTES));
        has_bias = T::kBias && g.bias != nullptr;
        has_gamma = T::kGamma && g.gamma != nullptr;
        has_resid = T::kResid && g.resid != nullptr;
        want_out = g.out != nullptr;
        // 256-bit accesses need 32-byte aligned row segments (chunk starts are multiples of 16 columns)
        auto aligned32 = [](const void* ptr, long long ld, int esz) {
            return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr) | (uintptr_t)(ld * esz)) & 31) == 0);
        };
        vec_ok_static = aligned32(g.out, g.ldo, OUT_ESZ) && aligned32(g.out2, g.ldo2, 2) &&
                        (!T::kResid || aligned32(g.resid, g.ldr, 4)) && (!T::kAux || aligned32(g.aux, g.ldaux, 2)) &&
                        ((g.so_h * OUT_ESZ) & 31) == 0 && ((g.so_b * OUT_ESZ) & 31) == 0;
    }

    // row state of tile (m_tile, n_tile, batch) for this lane
    __device__ __forceinline__ void make_row(int m_tile, int n_tile, int batch, EpiRow& R, int& n0) const {
        const long long ooff = (long long)(batch % g.nbatch_h) * g.so_h + (long long)(batch / g.nbatch_h) * g.so_b;
        const long long row = (long long)m_tile * GEMM_BM + quarter * 32 + lane;
        n0 = n_tile * BN + half * COLS_PER_WARP;
        R.ok = row < g.M;
        R.rs = 1.0f;
        R.out2 = nullptr; R.resid = nullptr; R.aux = nullptr;
        long long orow = row;  // output row
        if constexpr (EPI == EPI_TOKENS_F32) {
            long long bimg;
            int pp;
            if (g.pe_G > 0) {   // tile = (image, patch-row group): tile row r is patch group * pe_rows + r of the image
                const int r = quarter * 32 + lane;
                bimg = m_tile / g.pe_G;
                pp = (m_tile - (int)bimg * g.pe_G) * g.pe_rows + r;
                R.ok = r < g.pe_rows && pp < g.tok_n;
                if (!R.ok) pp = 0;
            } else {
                bimg = row / g.tok_n;
                pp = static_cast<int>(row - bimg * g.tok_n);
            }
            orow = bimg * g.tok_N + g.tok_T + pp;
            R.resid = g.resid + (long long)(g.tok_T + pp) * g.ldr + n0;  // pos_embed row
        } else if constexpr (EPI == EPI_RESID_F32) {
            if (has_resid) R.resid = g.resid + row * g.ldr + n0;
            if (g.out2 != nullptr) R.out2 = reinterpret_cast<__nv_bfloat16*>(g.out2) + row * g.ldo2 + n0;
            if (g.rowscale != nullptr && R.ok) R.rs = __ldg(g.rowscale + row / g.rows_per_sample);
        } else if constexpr (EPI == EPI_BIAS_GELU) {
            R.out2 = reinterpret_cast<__nv_bfloat16*>(g.out2) + row * g.ldo2 + n0;
        } else if constexpr (EPI == EPI_DGELU) {
            R.aux = g.aux + row * g.ldaux + n0;
        }
        R.out = want_out ? static_cast<void*>(reinterpret_cast<char*>(g.out) + (ooff + orow * g.ldo + n0) * OUT_ESZ)
                         : nullptr;
    }

    // all bulk stores issued by this warp have completed (call once before the kernel ends)
    __device__ __forceinline__ void finish() const {
        if (g.tma_out != 0 && EpiSmem<EPI>::kOutBufs > 0 && lane == 0) tma_store_wait<0>();
    }

    // L2 prefetch of the accumulator-independent operand rows (residual / GELU') of a future tile
    __device__ __forceinline__ void l2_prefetch(int m_tile, int n_tile, int batch) const {
        if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_DGELU) {
            if (g.dbg & 1) return;
            EpiRow R;
            int n0;
            make_row(m_tile, n_tile, batch, R, n0);
            const int nc = min(COLS_PER_WARP, g.N - n0) & ~7;  // 16-byte multiples for both element sizes
            if (!R.ok || nc <= 0) return;
            if constexpr (EPI == EPI_RESID_F32) {
                if (R.resid != nullptr) prefetch_l2_bulk(R.resid, nc * 4);
            } else {
                prefetch_l2_bulk(R.aux, nc * 2);
            }
        }
    }

    // Drain this warp's slab of the accumulator stage at `taddr` (lane / column offsets of this warp already applied).
    // `wait_acc()` blocks until the MMAs of the tile are complete; `release_acc()` is called once all TMEM reads of the
    // stage are done (by all lanes; the callee elects).
    // (the chunk math must be ONE basic block so that its 8 independent fp32x2 chains interleave: whether gelu' is
    // wanted is therefore a template parameter, not a branch inside the math)
    template <class WaitFn, class ReleaseFn>
    __device__ __forceinline__ void tile(int m_tile, int n_tile, int batch, uint32_t taddr, WaitFn wait_acc,
                                         ReleaseFn release_acc) {
        if (EPI != EPI_BIAS_GELU || want_out) tile_impl<true>(m_tile, n_tile, batch, taddr, wait_acc, release_acc);
        else tile_impl<false>(m_tile, n_tile, batch, taddr, wait_acc, release_acc);
    }
    template <bool WANT_DGELU, class WaitFn, class ReleaseFn>
    __device__ __forceinline__ void tile_impl(int m_tile, int n_tile, int batch, uint32_t taddr, WaitFn wait_acc,
                                              ReleaseFn release_acc) {
        EpiRow R;
        int n0;
        make_row(m_tile, n_tile, batch, R, n0);
        // stage bias / gamma of this warp's columns (zero / one padded beyond N)
        if constexpr (T::kBias) {
            if (lane * 4 < COLS_PER_WARP) {
                const int col = n0 + lane * 4;
                const float4 bv = (has_bias && col + 4 <= g.N) ? __ldg(reinterpret_cast<const float4*>(g.bias + col))
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(wvec + lane * 4) = bv;
                if constexpr (T::kGamma) {
                    const float4 gv = (has_gamma && col + 4 <= g.N)
                                          ? __ldg(reinterpret_cast<const float4*>(g.gamma + col))
                                          : make_float4(1.f, 1.f, 1.f, 1.f);
                    *reinterpret_cast<float4*>(wvec + COLS_PER_WARP + lane * 4) = gv;
                }
            }
            __syncwarp();
        }

        // operand ring: chunk c's residual / GELU' segment of this lane's row (vector path only)
        uint32_t opq[PF][T::kOpWords];
        auto chunk_is_vec = [&](int c) { return vec_ok_static && n0 + c * 16 + 16 <= g.N; };
        // TMA stores only when this warp's whole column slab is in range (boxes may span two chunks)
        const bool use_tma = g.tma_out != 0 && EpiSmem<EPI>::kOutBufs > 0 && n0 + COLS_PER_WARP <= g.N;
        auto fetch_chunk = [&](int c, uint32_t* dst) {
            if (!R.ok || !chunk_is_vec(c)) return;
            if (g.dbg & 4) return;
            if constexpr (T::kResid) {
                if (R.resid != nullptr) {
                    if (g.dbg & 2) {
                        ldg256_plain(R.resid + c * 16, dst);
                        ldg256_plain(R.resid + c * 16 + 8, dst + 8);
                    } else {
                    ldg256(R.resid + c * 16, dst);
                    ldg256(R.resid + c * 16 + 8, dst + 8);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[i] = 0u;
                }
            } else if constexpr (T::kAux) {
                ldg256(R.aux + c * 16, dst);
            }
        };
#pragma unroll
        for (int j = 0; j < PF; ++j) fetch_chunk(j, opq[j]);  // in flight while we wait for the MMAs

        wait_acc();
        tc_fence_after_sync();
        uint32_t racc[2][16];
        tmem_ld_32x32b_x16(taddr, racc[0]);
#pragma unroll 1
        for (int c0 = 0; c0 < NCHUNK; c0 += PF) {
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                const int c = c0 + j;
                const uint32_t* r = racc[j & 1];
                tmem_ld_wait();
                if (c + 1 < NCHUNK) {
                    // the accumulator chunk is fetched one chunk ahead of its use
                    tmem_ld_32x32b_x16(taddr + (c + 1) * 16, racc[(j + 1) & 1]);
                } else {
                    // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
                    tc_fence_before_sync();
                    __syncwarp();
                    release_acc();
                }
                const int col0 = n0 + c * 16;
                if (col0 >= g.N) continue;  // (warp-uniform)
                float cs[16];
                if (chunk_is_vec(c)) {
                    // ---- vector path: 16 columns, 256-bit global accesses
                    uint32_t ho[8], ho2[8], fo[16];
                    if constexpr (EPI == EPI_BIAS_GELU) {
                        // all 8 fp32x2 pairs of the chunk in lock step (see gelu_pairs)
                        float xv[16], gv[16], dv[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b4 = lds_f4(wvec_s + (c * 16 + 4 * q) * 4);
                            xv[4 * q] = __uint_as_float(r[4 * q]) + b4.x;
                            xv[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
                            xv[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z;
                            xv[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
                        }
                        gelu_pairs<8, WANT_DGELU>(xv, gv, dv);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            ho2[i] = pack_bf16(gv[2 * i], gv[2 * i + 1]);
                            if constexpr (WANT_DGELU) ho[i] = pack_bf16(dv[2 * i], dv[2 * i + 1]);
                        }
                    } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                     __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f), r4 = b4;
                        uint2 ax = make_uint2(0u, 0u);
                        if constexpr (T::kBias) b4 = lds_f4(wvec_s + (c * 16 + 4 * q) * 4);
                        if constexpr (T::kGamma) g4 = lds_f4(wvec_s + (COLS_PER_WARP + c * 16 + 4 * q) * 4);
                        if constexpr (T::kResid)
                            r4 = make_float4(__uint_as_float(opq[j][4 * q]), __uint_as_float(opq[j][4 * q + 1]),
                                             __uint_as_float(opq[j][4 * q + 2]), __uint_as_float(opq[j][4 * q + 3]));
                        if constexpr (T::kAux) ax = make_uint2(opq[j][2 * q], opq[j][2 * q + 1]);
                        float4 of;
                        uint2 oh, oh2;
                        const float4 w = epi_math4<EPI, WANT_DGELU>(v, b4, g4, r4, ax, R.rs, of, oh, oh2);
                        cs[4 * q] = w.x; cs[4 * q + 1] = w.y; cs[4 * q + 2] = w.z; cs[4 * q + 3] = w.w;
                        if constexpr (T::kOutF32) {
                            fo[4 * q] = __float_as_uint(of.x); fo[4 * q + 1] = __float_as_uint(of.y);
                            fo[4 * q + 2] = __float_as_uint(of.z); fo[4 * q + 3] = __float_as_uint(of.w);
                        } else {
                            ho[2 * q] = oh.x; ho[2 * q + 1] = oh.y;
                        }
                        if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32) {
                            ho2[2 * q] = oh2.x; ho2[2 * q + 1] = oh2.y;
                        }
                    }
                    }
                    if (use_tma) {
                        // ---- TMA store: stage this chunk in the warp's swizzled buffer; one lane issues the copy
                        if constexpr (EpiSmem<EPI>::kOutBufs > 0) {
                            constexpr int CHUNKS_PER_BOX = T::kOutF32 ? 1 : 2;   // 64-byte rows: 16 fp32 / 32 bf16 columns
                            const int sub = c % CHUNKS_PER_BOX;
                            if (sub == 0) {
                                // the previous TMA store out of this buffer must have finished reading it
                                if (lane == 0) tma_store_wait_read<0>();
                                __syncwarp();
                            }
                            const uint32_t rowb = stg_s + lane * 64;
                            const int sw = (lane >> 1) & 3;
                            if constexpr (T::kOutF32) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    sts_u4(rowb + ((q ^ sw) << 4), fo[4 * q], fo[4 * q + 1], fo[4 * q + 2], fo[4 * q + 3]);
                            } else {
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const int unit = sub * 2 + q;
                                    if (R.out != nullptr)
                                        sts_u4(rowb + ((unit ^ sw) << 4), ho[4 * q], ho[4 * q + 1], ho[4 * q + 2], ho[4 * q + 3]);
                                    if constexpr (EPI == EPI_BIAS_GELU)
                                        sts_u4(rowb + EPI_STG_BYTES + ((unit ^ sw) << 4), ho2[4 * q], ho2[4 * q + 1],
                                               ho2[4 * q + 2], ho2[4 * q + 3]);
                                }
                            }
                            if (sub == CHUNKS_PER_BOX - 1) {
                                fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0 && !(g.dbg & 8)) {
                                    const int col = n0 + (c - (CHUNKS_PER_BOX - 1)) * 16;
                                    const int row0 = m_tile * GEMM_BM + quarter * 32;  // rows >= M are clipped by TMA
                                    if (EPI != EPI_BIAS_GELU || want_out) tma_store_2d_s(tm_out, stg_s, col, row0);
                                    if constexpr (EPI == EPI_BIAS_GELU)
                                        tma_store_2d_s(tm_out2, stg_s + EPI_STG_BYTES, col, row0);
                                    tma_store_commit();
                                }
                            }
                        }
                        if constexpr (EPI == EPI_RESID_F32) {  // optional bf16 copy of the branch output (LayerScale models)
                            if (R.ok && R.out2 != nullptr) stg256(R.out2 + c * 16, ho2);
                        }
                    } else
                    if (R.ok && !(g.dbg & 8)) {
                        if constexpr (EPI == EPI_ATOMIC_F32) {
                            float* o = reinterpret_cast<float*>(R.out) + c * 16;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                red_add_v4_f32(o + 4 * q, __uint_as_float(fo[4 * q]), __uint_as_float(fo[4 * q + 1]),
                                               __uint_as_float(fo[4 * q + 2]), __uint_as_float(fo[4 * q + 3]));
                        } else if constexpr (T::kOutF32) {
                            float* o = reinterpret_cast<float*>(R.out) + c * 16;
                            stg256(o, fo);
                            stg256(o + 8, fo + 8);
                        } else {
                            if (R.out != nullptr) stg256(reinterpret_cast<__nv_bfloat16*>(R.out) + c * 16, ho);
                        }
                        if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32) {
                            if (R.out2 != nullptr) stg256(R.out2 + c * 16, ho2);
                        }
                    }
                    if (c + PF < NCHUNK) fetch_chunk(c + PF, opq[j]);  // refill the ring slot just consumed
                } else {
                    // ---- 4-column path: N tail or unaligned pointers
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cw = c * 16 + 4 * q;  // column within this warp's range
                        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (n0 + cw + 4 <= g.N) {
                            const float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                         __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f), r4 = b4;
                            uint2 ax = make_uint2(0u, 0u);
                            if constexpr (T::kBias) b4 = lds_f4(wvec_s + cw * 4);
                            if constexpr (T::kGamma) g4 = lds_f4(wvec_s + (COLS_PER_WARP + cw) * 4);
                            if (R.ok) {
                                if constexpr (T::kResid) {
                                    if (R.resid != nullptr) r4 = *reinterpret_cast<const float4*>(R.resid + cw);
                                }
                                if constexpr (T::kAux) ax = *reinterpret_cast<const uint2*>(R.aux + cw);
                            }
                            float4 of;
                            uint2 oh, oh2;
                            w = epi_math4<EPI, WANT_DGELU>(v, b4, g4, r4, ax, R.rs, of, oh, oh2);
                            if (R.ok) {
                                if constexpr (EPI == EPI_ATOMIC_F32) {
                                    red_add_v4_f32(reinterpret_cast<float*>(R.out) + cw, of.x, of.y, of.z, of.w);
                                } else if constexpr (T::kOutF32) {
                                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(R.out) + cw) = of;
                                } else {
                                    if (R.out != nullptr)
                                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(R.out) + cw) = oh;
                                }
                                if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32) {
                                    if (R.out2 != nullptr) *reinterpret_cast<uint2*>(R.out2 + cw) = oh2;
                                }
                            }
                        }
                        cs[4 * q] = w.x; cs[4 * q + 1] = w.y; cs[4 * q + 2] = w.z; cs[4 * q + 3] = w.w;
                    }
                }
                if constexpr (T::kColsum) {
                    if (g.colsum != nullptr) {
                        if (!R.ok) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) cs[i] = 0.f;
                        }
                        const float tot = warp_colsum16(cs, lane);
                        const int col = col0 + warp_colsum16_col(lane);
                        if ((lane & 1) == 0 && col < g.N) atomicAdd(g.colsum + col, tot);
                    }
                }
            }
        }
    }
};

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const GemmArgs g) {
    using Cfg = GemmCfg<BN, EPI>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* svec = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);  // per-warp bias | gamma
    uint8_t* sstg = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Cfg::VEC_BYTES) + 1023) &
        ~uintptr_t(1023));  // TMA-store staging, 2 KB per epilogue warp and output

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

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
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int num_units = g.num_m_tiles * g.num_n_tiles * g.splits * g.nbatch_h * g.nbatch_b;
    const bool batched = g.nbatch_h * g.nbatch_b > 1;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int n_tile = u % g.num_n_tiles;
                const int rest = u / g.num_n_tiles;
                const int split = rest % g.splits;
                const int mb = rest / g.splits;
                const int m_tile = mb % g.num_m_tiles;
                const int batch = mb / g.num_m_tiles;
                const int bh = batch % g.nbatch_h, bb = batch / g.nbatch_h;
                const int m0 = m_tile * GEMM_BM, n0 = n_tile * BN;
                const int kb0 = split * g.kblocks_per_split;
                const int kb1 = min(kb0 + g.kblocks_per_split, g.num_kblocks);
                if (!batched) {
                    // plain GEMM: rank-2 tensor maps (rank-4 boxes cost 20-45% on MN-major operands)
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_role(&empty_bar[stage], phase ^ 1, g.dbg);
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        uint8_t* sa = smem_a + stage * Cfg::A_STAGE_BYTES;
                        uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
                        const int k0 = kb * GEMM_BK;
                        if constexpr (!A_MN) {
                            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);  // box {64 k, 128 m}
                        } else {
#pragma unroll
                            for (int j = 0; j < GEMM_BM / 64; ++j)  // box {64 m, 64 k} per 64-wide M chunk
                                tma_load_2d(sa + j * (GEMM_BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, k0);
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);  // box {64 k, BN n}
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                tma_load_2d(sb + j * (GEMM_BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, k0);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                } else {
                    // batched: rank-4 maps {inner, row, batch_h, batch_b}; the three outer coordinates are permuted into
                    // tensor-map dimension order (dims are sorted by stride on the host)
                    auto load_tile = [&](void* dst, const CUtensorMap* tm, const int* perm, int c0, int row) {
                        const int lc[3] = {row, bh, bb};
                        tma_load_4d(dst, tm, &full_bar[stage], c0, lc[perm[0]], lc[perm[1]], lc[perm[2]]);
                    };
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_role(&empty_bar[stage], phase ^ 1, g.dbg);
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        uint8_t* sa = smem_a + stage * Cfg::A_STAGE_BYTES;
                        uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
                        const int k0 = kb * GEMM_BK;
                        if constexpr (!A_MN) {
                            load_tile(sa, &tmA, g.a_perm, k0, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < GEMM_BM / 64; ++j)
                                load_tile(sa + j * (GEMM_BK * 128), &tmA, g.a_perm, m0 + 64 * j, k0);
                        }
                        if constexpr (!B_MN) {
                            load_tile(sb, &tmB, g.b_perm, k0, n0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                load_tile(sb + j * (GEMM_BK * 128), &tmB, g.b_perm, n0 + 64 * j, k0);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            constexpr uint32_t A_LBO = A_MN ? GEMM_BK * 128 : 0, B_LBO = B_MN ? GEMM_BK * 128 : 0;
            // per-UMMA K advance (16 elements): +32 B inside the 128 B swizzle row (K-major) / +16 rows of 128 B (MN-major)
            constexpr uint32_t A_KSTEP = (A_MN ? 2048u : 32u) >> 4, B_KSTEP = (B_MN ? 2048u : 32u) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int rest = u / g.num_n_tiles;
                const int split = rest % g.splits;
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
                        umma_bf16(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= GEMM_EPI_WARP0) {
        // ===================== epilogue =====================
        const int ew = warp - GEMM_EPI_WARP0;
        EpilogueWarp<BN, EPI> epi(g, svec, sstg, &tmO, &tmO2, ew, warp, lane);
        int as = 0;
        uint32_t aphase = 0;
        auto decode = [&](int un, int& m_tile, int& n_tile, int& batch) {
            n_tile = un % g.num_n_tiles;
            const int mb = (un / g.num_n_tiles) / g.splits;
            m_tile = mb % g.num_m_tiles;
            batch = mb / g.num_m_tiles;
        };
        int m_tile, n_tile, batch;
        decode(blockIdx.x, m_tile, n_tile, batch);
        epi.l2_prefetch(m_tile, n_tile, batch);
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            if (u + (int)gridDim.x < num_units) {  // a whole tile of lead time
                decode(u + gridDim.x, m_tile, n_tile, batch);
                epi.l2_prefetch(m_tile, n_tile, batch);
            }
            decode(u, m_tile, n_tile, batch);
            const uint32_t taddr =
                tmem_base + (uint32_t((warp & 3) * 32) << 16) + as * BN + (ew >> 2) * (BN / 2);
            epi.tile(m_tile, n_tile, batch, taddr, [&]() { mbar_wait(&tfull_bar[as], aphase); },
                     [&]() { if (lane == 0) mbar_arrive(&tempty_bar[as]); });
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        epi.finish();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace vitk
