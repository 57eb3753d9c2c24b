// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (128B swizzle) -> shared-memory ring -> tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulators in TMEM,
//   double-buffered) -> tcgen05.ld epilogue with fused bias / GELU / residual / LayerScale / GELU-grad / split-K.
//
//   D[M,N] = A[M,K] * B[N,K]^T            (A and B independently K-major or MN-major in global memory)
//
// Replaces on the reference path: nn.Linear forward (models/cait.py:99,102,113,126; timm/DINO Attention.qkv/proj,
// Mlp.fc1/fc2 -- in-repo witness models/swin.py:24-30) and the autograd dgrad / wgrad of the same Linears.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4..11 = epilogue (two warps per TMEM lane quarter, each taking half of the BN columns).
#pragma once
#include "common.cuh"

namespace vitk {

enum GemmEpilogue : int {
    EPI_STORE_BF16 = 0,  // out_bf16 = acc (+bias)
    EPI_BIAS_GELU = 1,   // pre = acc + bias ; out_bf16 = pre ; out2_bf16 = gelu(pre)
    EPI_RESID_F32 = 2,   // v = acc + bias ; [out2_bf16 = v] ; out_f32 = resid + gamma * v   (gamma optional)
    EPI_DGELU = 3,       // out_bf16 = acc * gelu'(aux_bf16)
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
    float* colsum;  // optional fp32 [N]: += column sums of the values stored to `out` (bias gradient of the next Linear)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN> struct GemmCfg {
    static constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    static constexpr int B_STAGE_BYTES = BN * GEMM_BK * 2;       // 32 KB (BN=256) / 16 KB (BN=128)
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 256;
    static constexpr int VEC_BYTES = 2 * 2 * BN * 4;  // double-buffered bias[BN] and gamma[BN] for the epilogue
    static constexpr int STG_BYTES = GEMM_EPI_WARPS * 32 * 16 * 4;  // per-warp [32 rows x 16 fp32] transpose buffer
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + VEC_BYTES + STG_BYTES + 1024;  // +1 KB: align
};

// ----------------------------------------------------------------------------------------------------------------
// Epilogue math on 4 consecutive columns of one row, executed in the COALESCED domain: after the accumulator chunk has
// been transposed through shared memory, the 8 lanes lane/8.. of a warp own 4 columns each of the same row, so that
// every global load/store instruction touches whole 32-byte sectors of a few rows.
// ----------------------------------------------------------------------------------------------------------------
struct EpiOperand {  // operands that do not depend on the accumulator; fetched before the accumulator is waited on
    float4 r;        // residual (EPI_RESID_F32) / positional embedding (EPI_TOKENS_F32)
    uint2 aux;       // 4 bf16 pre-activations (EPI_DGELU)
};

template <int EPI>
__device__ __forceinline__ void epi_fetch(const GemmArgs& g, EpiOperand& op, long long row, int col) {
    if constexpr (EPI == EPI_RESID_F32) {
        if (g.resid != nullptr) op.r = *reinterpret_cast<const float4*>(g.resid + row * g.ldr + col);
    } else if constexpr (EPI == EPI_DGELU) {
        op.aux = *reinterpret_cast<const uint2*>(g.aux + row * g.ldaux + col);
    } else if constexpr (EPI == EPI_TOKENS_F32) {
        const long long bimg = row / g.tok_n;
        const int p = static_cast<int>(row - bimg * g.tok_n);
        op.r = __ldg(reinterpret_cast<const float4*>(g.resid + (long long)(g.tok_T + p) * g.ldr + col));
    }
}

template <int EPI>
__device__ __forceinline__ float4 epi_apply(const GemmArgs& g, float4 v, long long row, int col, const float4 b,
                                            const float4 gm, const EpiOperand& op, long long ooff) {
    // b / gm: bias and LayerScale gamma of these 4 columns (zeros / ones when absent)
    if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32 ||
                  EPI == EPI_TOKENS_F32) {
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if constexpr (EPI == EPI_STORE_BF16) {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + ooff + row * g.ldo + col) =
            make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else if constexpr (EPI == EPI_BIAS_GELU) {
        if (g.out != nullptr)  // pre-activation is only needed for backward
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + row * g.ldo + col) =
                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + row * g.ldo2 + col) =
            make_uint2(pack_bf16(gelu_f(v.x), gelu_f(v.y)), pack_bf16(gelu_f(v.z), gelu_f(v.w)));
    } else if constexpr (EPI == EPI_RESID_F32) {
        if (g.out2 != nullptr)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + row * g.ldo2 + col) =
                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        v.x *= gm.x; v.y *= gm.y; v.z *= gm.z; v.w *= gm.w;
        if (g.rowscale != nullptr) {
            const float rs = __ldg(g.rowscale + row / g.rows_per_sample);
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
        }
        if (g.resid != nullptr) { v.x += op.r.x; v.y += op.r.y; v.z += op.r.z; v.w += op.r.w; }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + row * g.ldo + col) = v;
    } else if constexpr (EPI == EPI_DGELU) {
        v.x *= gelu_grad_f(bf16_lo(op.aux.x)); v.y *= gelu_grad_f(bf16_hi(op.aux.x));
        v.z *= gelu_grad_f(bf16_lo(op.aux.y)); v.w *= gelu_grad_f(bf16_hi(op.aux.y));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + row * g.ldo + col) =
            make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else if constexpr (EPI == EPI_ATOMIC_F32) {
        red_add_v4_f32(reinterpret_cast<float*>(g.out) + row * g.ldo + col, v.x, v.y, v.z, v.w);
    } else if constexpr (EPI == EPI_STORE_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + ooff + row * g.ldo + col) = v;
    } else if constexpr (EPI == EPI_TOKENS_F32) {
        const long long bimg = row / g.tok_n;
        const int p = static_cast<int>(row - bimg * g.tok_n);
        v.x += op.r.x; v.y += op.r.y; v.z += op.r.z; v.w += op.r.w;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + (bimg * g.tok_N + g.tok_T + p) * g.ldo + col) = v;
    }
    return v;  // the value written to `out` (for STORE_* / DGELU / ATOMIC; used by the optional column-sum reduction)
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    using Cfg = GemmCfg<BN>;
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
    float* svec = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);  // [2][bias BN | gamma BN]
    float* sstage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Cfg::VEC_BYTES);

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
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
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
                        mbar_wait(&empty_bar[stage], phase ^ 1);
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
                        mbar_wait(&empty_bar[stage], phase ^ 1);
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
                mbar_wait(&tempty_bar[as], aphase ^ 1);
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
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = ew >> 2;      // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        int as = 0;
        uint32_t aphase = 0;
        const int et = threadIdx.x - GEMM_EPI_WARP0 * 32;  // 0..255 within the epilogue warps
        constexpr bool kUsesVec = (EPI == EPI_STORE_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_RESID_F32 ||
                                   EPI == EPI_STORE_F32 || EPI == EPI_TOKENS_F32);
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int n_tile = u % g.num_n_tiles;
            const int rest = u / g.num_n_tiles;
            const int mb = rest / g.splits;
            const int m_tile = mb % g.num_m_tiles;
            const int batch = mb / g.num_m_tiles;
            const long long ooff = (long long)(batch % g.nbatch_h) * g.so_h + (long long)(batch / g.nbatch_h) * g.so_b;
            const bool has_bias = kUsesVec && g.bias != nullptr;
            const bool has_gamma = (EPI == EPI_RESID_F32) && g.gamma != nullptr;
            // coalesced-domain coordinates of this lane: 8 rows per pass, 4 lanes x 4 columns per row
            const long long row_base = (long long)m_tile * GEMM_BM + quarter * 32;
            const int n0 = n_tile * BN + half * COLS_PER_WARP;
            const int sub_row = lane >> 2, sub_col = (lane & 3) * 4;
            const uint32_t stg = smem_u32(sstage + ew * (32 * 16));
            constexpr int NCHUNK = COLS_PER_WARP / 16;
            EpiOperand nxt[4];
            auto fetch_chunk = [&](int c) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const long long rr = row_base + it * 8 + sub_row;
                    const int col = n0 + c * 16 + sub_col;
                    if (rr < g.M && col + 4 <= g.N) epi_fetch<EPI>(g, nxt[it], rr, col);
                }
            };
            fetch_chunk(0);  // operands that do not depend on the accumulator: in flight while we wait for the MMAs
            // bias (and LayerScale gamma) of this lane's 4 columns in each of the NCHUNK chunks: plain read-only loads
            // issued before the accumulator wait. (A shared-memory staging + block barrier per tile made all eight
            // epilogue warps rendezvous and cost ~15% on the K=768 forward GEMMs.)
            float4 bias4[NCHUNK], gam4[(EPI == EPI_RESID_F32) ? NCHUNK : 1];
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int col = n0 + c * 16 + sub_col;
                bias4[c] = (has_bias && col + 4 <= g.N) ? __ldg(reinterpret_cast<const float4*>(g.bias + col))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (EPI == EPI_RESID_F32)
                    gam4[c] = (has_gamma && col + 4 <= g.N) ? __ldg(reinterpret_cast<const float4*>(g.gamma + col))
                                                            : make_float4(1.f, 1.f, 1.f, 1.f);
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN + half * COLS_PER_WARP;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(taddr + c * 16, r);
                tmem_ld_wait();
                if (c == NCHUNK - 1) {
                    // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[as]);
                }
                // transpose through shared memory: TMEM domain (lane == row) -> coalesced domain. 16-byte units are
                // XOR-swizzled with (row >> 1) so that both the row-wise writes and the 8-row reads are conflict free.
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    sts_u4(stg + (lane * 16 + ((u ^ (lane >> 1)) & 3) * 4) * 4, r[u * 4 + 0], r[u * 4 + 1], r[u * 4 + 2],
                           r[u * 4 + 3]);
                __syncwarp();
                // batch every shared-memory read of the chunk before the math / global stores (2 warps per scheduler
                // cannot hide a load-use chain per element)
                float4 v[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int lr = it * 8 + sub_row;
                    v[it] = lds_f4(stg + (lr * 16 + (((lane & 3) ^ (lr >> 1)) & 3) * 4) * 4);
                }
                const float4 b4 = bias4[c];
                const float4 g4 = (EPI == EPI_RESID_F32) ? gam4[(EPI == EPI_RESID_F32) ? c : 0]
                                                         : make_float4(1.f, 1.f, 1.f, 1.f);
                EpiOperand cur[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) cur[it] = nxt[it];
                if (c + 1 < NCHUNK) fetch_chunk(c + 1);
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const long long rr = row_base + it * 8 + sub_row;
                    const int col = n0 + c * 16 + sub_col;
                    if (rr < g.M && col + 4 <= g.N) {
                        const float4 w = epi_apply<EPI>(g, v[it], rr, col, b4, g4, cur[it], ooff);
                        cs.x += w.x; cs.y += w.y; cs.z += w.z; cs.w += w.w;
                    }
                }
                if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_DGELU) {
                    if (g.colsum != nullptr) {
                        // reduce over the warp's 32 rows: lanes with equal (lane & 3) own the same 4 columns
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
                            cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
                            cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
                        }
                        const int col = n0 + c * 16 + sub_col;
                        if (lane < 4 && col + 4 <= g.N) red_add_v4_f32(g.colsum + col, cs.x, cs.y, cs.z, cs.w);
                    }
                }
                __syncwarp();
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace vitk
