// PatchEmbed as an im2col-FREE patch GEMM for sm_100a: the A operand is gathered by TMA straight from the NCHW image.
//
//   reference: Conv2d(C, D, kernel=P, stride=P)(x).flatten(2).transpose(1, 2)  (DINO / timm PatchEmbed, call site
//   models/vision_all.py:156,161-167; in-repo witness models/swin.py:434-445), then cls / pos assembly
//   (models/cait.py:229-234, models/deit.py:35-43).
//
// The image is described to the TMA unit as a rank-5 tensor {px: P, pc: W/P, py: P, pr: H/P, cb: C*B} (strides 1, P, W,
// P*W, H*W elements). One box {P, W/P, 1, RPT, 1} lands in shared memory as RPT*W/P consecutive rows -- one row per
// patch, P contiguous pixels of image row py of channel c -- which is exactly a K-major UMMA operand tile with rows of
// RB = P * sizeof(T) bytes (SWIZZLE_32B / 64B / 128B for RB = 32 / 64 / 128). A k-block of the GEMM is therefore
// (channel c, image row py) and K = C*P*P is walked in C*P k-blocks of P elements; no [B*n, C*P*P] patch matrix exists
// in HBM, neither in forward nor in backward.
//
//   fp32 images  -> tcgen05.mma kind::tf32 on the fp32 pixels and the fp32 master weight (no cast pass at all)
//   bf16 images  -> tcgen05.mma kind::f16 on bf16 pixels and the bf16 weight copy
//
// forward : out[b, T + p, :] = sum_k patch[b,p,k] W[:,k] + bias + pos[T + p]      (persistent, GEMM epilogue engine)
// backward: dW[d, k] += sum_{b,p} dY[b, T+p, d] patch[b,p,k]                       (both operands MN-major, split over
//           images, red.global.add into the fp32 gradient); dY is read in place from the [B, N, D] token gradient.
#include "gemm.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

template <typename T> struct PeElem;
template <> struct PeElem<float> {
    static constexpr int ESZ = 4, UMMA_K = 8, FMT = 2;  // TF32
    static constexpr CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};
template <> struct PeElem<__nv_bfloat16> {
    static constexpr int ESZ = 2, UMMA_K = 16, FMT = 1;  // BF16
    static constexpr CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};

// instruction descriptor: c_format F32, a/b format FMT (1 = BF16, 2 = TF32), major bits, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t pe_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// shared-memory matrix descriptor with an explicit swizzle mode (layout type 2 = 128B, 4 = 64B, 6 = 32B)
__device__ __forceinline__ uint64_t pe_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}
// UMMA layout type of a tile whose rows are rb bytes: 128B / 64B / 32B swizzle; 16-byte rows use the unswizzled
// ("interleaved") canonical layout, where 8 consecutive 16-byte rows form one core matrix
__device__ __forceinline__ uint32_t pe_layout_of(int rb) { return rb == 128 ? 2u : (rb == 64 ? 4u : (rb == 32 ? 6u : 0u)); }

template <typename T>
__device__ __forceinline__ void pe_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if constexpr (sizeof(T) == 4) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        umma_bf16(tmem_d, adesc, bdesc, idesc, acc);
    }
}

struct PeGeom {
    int B, C, P, Hp, Wp;  // image grid
    int RB;               // bytes of one patch row in shared memory = P * sizeof(T)
    int KPS;              // k-blocks (image rows py) per 128-byte pipeline stage = 128 / RB
    int rpt, G;           // patch rows per M tile, M tiles per image
    int rows;             // rpt * Wp: rows of an M tile that carry patches
    int num_stages_k;     // C * P / KPS
};

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
constexpr int PE_BN = 256;
constexpr int PE_STAGES = 4;
constexpr int PE_A_BYTES = GEMM_BM * 128;       // 16 KB: KPS sub-tiles of [128 rows x RB]
constexpr int PE_B_BYTES = PE_BN * 128;         // 32 KB: [256 weight rows x 128 B of k]
constexpr int PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_VEC_BYTES = GEMM_EPI_WARPS * 2 * (PE_BN / 2) * 4;
constexpr int PE_FWD_SMEM = PE_STAGES * PE_STAGE_BYTES + 256 + PE_VEC_BYTES + 2048;

template <typename T>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
patch_embed_fwd_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmW,
                       const GemmArgs g, const PeGeom pe) {
    using E = PeElem<T>;
    constexpr int STAGES = PE_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * PE_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * PE_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* svec = reinterpret_cast<float*>(smem + STAGES * PE_STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmImg);
        tma_prefetch_desc(&tmW);
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
    if (warp == 1) tmem_alloc<2 * PE_BN>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int num_units = g.num_m_tiles * g.num_n_tiles;
    const int a_sub_bytes = GEMM_BM * pe.RB;   // one [128 x RB] sub-tile
    const int py_per_c = pe.P;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(pe.KPS * pe.rows * pe.RB + PE_B_BYTES);
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int n_tile = u % g.num_n_tiles, m_tile = u / g.num_n_tiles;
                const int b = m_tile / pe.G, pr0 = (m_tile - b * pe.G) * pe.rpt;
                for (int ks = 0; ks < pe.num_stages_k; ++ks) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], tx);
                    uint8_t* sa = smem_a + stage * PE_A_BYTES;
                    uint8_t* sb = smem_b + stage * PE_B_BYTES;
                    const int kb0 = ks * pe.KPS;                      // first (c, py) k-block of this stage
                    for (int j = 0; j < pe.KPS; ++j) {
                        const int kb = kb0 + j, c = kb / py_per_c, py = kb - c * py_per_c;
                        tma_load_5d(sa + j * a_sub_bytes, &tmImg, &full_bar[stage], 0, 0, py, pr0, b * pe.C + c);
                    }
                    // weight [D, C*P*P], k = (c, py, px): KPS consecutive k-blocks are 128 contiguous bytes
                    tma_load_2d(sb, &tmW, &full_bar[stage], kb0 * pe.P, n_tile * PE_BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = pe_idesc(E::FMT, GEMM_BM, PE_BN, 0, 0);
            const uint32_t a_layout = pe_layout_of(pe.RB);
            const int mma_per_sub = pe.RB >= 32 ? pe.RB / 32 : 1;   // each MMA consumes 32 bytes of k per row
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + as * PE_BN;
                for (int ks = 0; ks < pe.num_stages_k; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem_a + stage * PE_A_BYTES);
                    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * PE_B_BYTES), 0, 1024);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint64_t adesc;
                        if (pe.RB >= 32) {
                            const int sub = i / mma_per_sub, within = i - sub * mma_per_sub;
                            adesc = pe_smem_desc(sa + sub * a_sub_bytes + within * 32, 0, 8 * pe.RB, a_layout);
                        } else {
                            // 16-byte patch rows: the two 16-byte K chunks of one MMA are image rows py, py+1 = two
                            // sub-tiles (LBO apart); 8 patches = one 128-byte core matrix (SBO apart)
                            adesc = pe_smem_desc(sa + 2 * i * a_sub_bytes, (uint32_t)a_sub_bytes, 128, 0u);
                        }
                        pe_umma<T>(tmem_d, adesc, bdesc + i * 2, idesc, (ks > 0 || i > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= GEMM_EPI_WARP0) {
        const int ew = warp - GEMM_EPI_WARP0;
        EpilogueWarp<PE_BN, EPI_TOKENS_F32> epi(g, svec, nullptr, nullptr, nullptr, ew, warp, lane);
        int as = 0;
        uint32_t aphase = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int n_tile = u % g.num_n_tiles, m_tile = u / g.num_n_tiles;
            const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + as * PE_BN + (ew >> 2) * (PE_BN / 2);
            epi.tile(m_tile, n_tile, 0, taddr, [&]() { mbar_wait(&tfull_bar[as], aphase); },
                     [&]() { if (lane == 0) mbar_arrive(&tempty_bar[as]); });
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<2 * PE_BN>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: dW[d, k] += sum over patches of dY[patch, d] * pixel[patch, k]
//   A' = dY^T : MN-major, [patch rows x 128 B of d] chunks (SWIZZLE_128B), 128 d per unit
//   B' = patch^T: MN-major, [patch rows x RB] chunks, one per (c, py) k-block, NCH chunks = N / P columns of dW per unit
// A pipeline stage holds WG_ROWS = 64 patch rows; a box brings rpt_b * Wp <= 64 of them, the tail rows stay zero
// (zeroed once: TMA never writes them), so the K loop can run in whole UMMA_K steps.
// ------------------------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 64;
constexpr int WG_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct PeWgArgs {
    int D, K;             // dW is [D, K] fp32, K = C*P*P
    int NCH;              // (c, py) k-blocks per unit; unit width N = NCH * P columns
    int kgroups;          // C*P / NCH
    int d_tiles;          // ceil(D / 128)
    int rpt_b, Gb;        // patch rows per box, boxes per image
    int rows_b;           // rpt_b * Wp
    int ksteps;           // ceil(rows_b / UMMA_K)
    int total_it, it_per_split, splits;
    int tok_T;            // first patch token in the dY token axis
    float* dW;
};

template <typename T>
__global__ void __launch_bounds__(WG_THREADS, 1)
patch_embed_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmImg,
                         const PeWgArgs w, const PeGeom pe, const int stages) {
    using E = PeElem<T>;
    constexpr int MCH = 128 * E::ESZ / 128;            // 128-byte chunks of d per unit: 4 (fp32) / 2 (bf16)
    constexpr int DCH = 128 / E::ESZ;                  // d elements per chunk
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = MCH * WG_ROWS * 128;
    const int b_chunk = WG_ROWS * pe.RB;
    const int b_bytes = ((w.NCH * b_chunk) + 1023) & ~1023;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tfull_bar = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // zero the stage buffers once: rows >= rows_b of every chunk are never written by TMA and must read as zeros
    for (int i = threadIdx.x; i < stages * stage_bytes / 16; i += WG_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmImg);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int u = blockIdx.x;
    const int d_tile = u % w.d_tiles;
    const int kg = (u / w.d_tiles) % w.kgroups;
    const int split = u / (w.d_tiles * w.kgroups);
    const int it0 = split * w.it_per_split;
    const int it1 = min(it0 + w.it_per_split, w.total_it);
    const int N = w.NCH * pe.P;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(w.rows_b * (MCH * 128 + w.NCH * pe.RB));
            for (int it = it0; it < it1; ++it) {
                const int b = it / w.Gb, pr0 = (it - b * w.Gb) * w.rpt_b;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], tx);
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
#pragma unroll
                for (int j = 0; j < MCH; ++j)  // dY[b, T + pr0*Wp ..., d_tile*128 + j*DCH ...]: box {DCH, rows_b, 1}
                    tma_load_3d(sa + j * (WG_ROWS * 128), &tmDy, &full_bar[stage], d_tile * 128 + j * DCH,
                                w.tok_T + pr0 * pe.Wp, b);
                for (int j = 0; j < w.NCH; ++j) {
                    const int kb = kg * w.NCH + j, c = kb / pe.P, py = kb - c * pe.P;
                    tma_load_5d(sb + j * b_chunk, &tmImg, &full_bar[stage], 0, 0, py, pr0, b * pe.C + c);
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = pe_idesc(E::FMT, 128, (uint32_t)N, 1, 1);
            const uint32_t b_layout = pe_layout_of(pe.RB);
            const uint32_t a_kstep = E::UMMA_K * 128, b_kstep = (uint32_t)(E::UMMA_K * pe.RB);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                const uint32_t sb = sa + a_bytes;
                for (int k = 0; k < w.ksteps; ++k) {
                    // MN-major: LBO = distance between MN chunks, SBO = distance between 8-row k groups
                    const uint64_t adesc = pe_smem_desc(sa + k * a_kstep, WG_ROWS * 128, 8 * 128, 2u);
                    // (unswizzled 16-byte rows: the descriptor's two offsets swap roles, see make_umma_desc<Major::MN>)
                    const uint64_t bdesc = pe.RB >= 32
                                               ? pe_smem_desc(sb + k * b_kstep, (uint32_t)b_chunk, 8 * pe.RB, b_layout)
                                               : pe_smem_desc(sb + k * b_kstep, 8 * pe.RB, (uint32_t)b_chunk, 0u);
                    pe_umma<T>(tmem_base, adesc, bdesc, idesc, (it > it0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull_bar);
        }
    } else {
        // epilogue: lane = d row of the unit, 16 accumulator columns (= k of dW) per step
        const int q = warp & 3;
        const int d = d_tile * 128 + q * 32 + lane;
        if (it1 > it0) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
            float* orow = w.dW + (long long)d * w.K + (long long)kg * N;
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(taddr + c0, r);
                tmem_ld_wait();
                if (d < w.D) {
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        red_add_v4_f32(orow + c0 + 4 * v, __uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                       __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<256>(tmem_base);
    }
}

// u8 [B,C,H,W] -> bf16 (x / 255 - mean[c]) / std[c]: ToTensor + Normalize (utils_datasets.py:573-580) on the device
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ stdv, long long total16, int C, long long hw) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (long long)gridDim.x * blockDim.x) {
        const long long e0 = i * 16;
        const int c = static_cast<int>((e0 / hw) % C);
        const float s = 1.0f / (255.0f * __ldg(stdv + c)), o = -__ldg(mean + c) / __ldg(stdv + c);
        const uint4 v = ld_nc_v4(x + e0);
        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
        uint32_t out[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float f0 = fmaf((float)(wds[k] & 0xff), s, o), f1 = fmaf((float)((wds[k] >> 8) & 0xff), s, o);
            const float f2 = fmaf((float)((wds[k] >> 16) & 0xff), s, o), f3 = fmaf((float)(wds[k] >> 24), s, o);
            out[2 * k] = pack_bf16(f0, f1);
            out[2 * k + 1] = pack_bf16(f2, f3);
        }
        st_v4(y + e0, make_uint4(out[0], out[1], out[2], out[3]));
        st_v4(y + e0 + 8, make_uint4(out[4], out[5], out[6], out[7]));
    }
}

}  // namespace vitk

using namespace vitk;

// rank-5 image map {px, pc, py, pr, cb}; box {P, Wp, 1, rpt, 1}
static int make_img_tmap(CUtensorMap* tm, const void* img, int esz, CUtensorMapDataType dt, int B, int C, int H, int W,
                         int P, int rpt) {
    const uint64_t Hp = H / P, Wp = W / P;
    uint64_t dims[5] = {(uint64_t)P, Wp, (uint64_t)P, Hp, (uint64_t)B * C};
    uint64_t strides[4] = {(uint64_t)P * esz, (uint64_t)W * esz, (uint64_t)P * W * esz, (uint64_t)H * W * esz};
    uint32_t box[5] = {(uint32_t)P, (uint32_t)Wp, 1u, (uint32_t)rpt, 1u};
    const int rb = P * esz;
    const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : (rb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    return make_tmap(tm, img, 5, dims, strides, box, sw, dt);
}

static bool pe_geom(PeGeom& pe, int B, int C, int H, int W, int P, int esz, int max_rows) {
    if (B <= 0 || C <= 0 || P <= 0 || H % P || W % P) return false;
    const int rb = P * esz;
    if (!(rb == 16 || rb == 32 || rb == 64 || rb == 128)) return false;
    pe.B = B; pe.C = C; pe.P = P; pe.Hp = H / P; pe.Wp = W / P;
    if (pe.Wp > max_rows || pe.Wp > 256 || pe.Hp > 256) return false;
    pe.RB = rb; pe.KPS = 128 / rb;
    if ((C * P) % pe.KPS) return false;
    pe.rpt = max_rows / pe.Wp;
    if (pe.rpt > pe.Hp) pe.rpt = pe.Hp;
    pe.G = (pe.Hp + pe.rpt - 1) / pe.rpt;
    pe.rows = pe.rpt * pe.Wp;
    pe.num_stages_k = C * P / pe.KPS;
    if (((long long)W * esz) % 16 || ((long long)H * W * esz) % 16) return false;
    return true;
}

template <typename T>
static int pe_fwd_launch(const void* img, const void* weight, const float* bias, const float* pos, long long ldpos,
                         float* out, int B, int C, int H, int W, int P, int D, int tok_N, int tok_T, cudaStream_t st) {
    using E = PeElem<T>;
    PeGeom pe;
    if (!pe_geom(pe, B, C, H, W, P, E::ESZ, GEMM_BM)) return VITK_ERR_UNSUPPORTED;
    const int n = pe.Hp * pe.Wp, K = C * P * P;
    if (tok_N < n + tok_T || (D % 4) != 0 || ((long long)K * E::ESZ) % 16) return VITK_ERR_ARG;
    CUtensorMap tmImg, tmW;
    if (make_img_tmap(&tmImg, img, E::ESZ, E::DT, B, C, H, W, P, pe.rpt)) return VITK_ERR_TMAP;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)D};
        uint64_t strides[1] = {(uint64_t)K * E::ESZ};
        uint32_t box[2] = {(uint32_t)(128 / E::ESZ), (uint32_t)PE_BN};
        if (make_tmap(&tmW, weight, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, E::DT)) return VITK_ERR_TMAP;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = B * pe.G * GEMM_BM; g.N = D; g.K = K;
    g.num_m_tiles = B * pe.G;
    g.num_n_tiles = (D + PE_BN - 1) / PE_BN;
    g.splits = 1; g.num_kblocks = pe.num_stages_k; g.kblocks_per_split = pe.num_stages_k;
    g.bias = bias; g.resid = pos; g.ldr = ldpos;
    g.out = out; g.ldo = D;
    g.tok_n = n; g.tok_N = tok_N; g.tok_T = tok_T;
    g.nbatch_h = g.nbatch_b = 1;
    g.pe_G = pe.G; g.pe_rows = pe.rows;
    auto kern = patch_embed_fwd_kernel<T>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_FWD_SMEM) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr_set = true;
    }
    const int units = g.num_m_tiles * g.num_n_tiles;
    const int grid = units < sm_count() ? units : sm_count();
    kern<<<grid, GEMM_THREADS, PE_FWD_SMEM, st>>>(tmImg, tmW, g, pe);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

template <typename T>
static int pe_wgrad_launch(const void* img, const void* dy, int dy_tok_N, int dy_tok_T, float* dW, int B, int C, int H,
                           int W, int P, int D, cudaStream_t st) {
    using E = PeElem<T>;
    PeGeom pe;
    if (!pe_geom(pe, B, C, H, W, P, E::ESZ, WG_ROWS)) return VITK_ERR_UNSUPPORTED;
    const int n = pe.Hp * pe.Wp;
    if (dy_tok_N < n + dy_tok_T || ((long long)D * E::ESZ) % 16) return VITK_ERR_ARG;
    PeWgArgs w;
    w.D = D; w.K = C * P * P;
    // unit width: the largest number of (c, py) k-blocks that divides C*P and keeps N = NCH*P <= 256
    const int kblocks = C * P;
    int nch = 256 / P;
    while (nch > 1 && kblocks % nch) --nch;
    if ((nch * P) % 16) return VITK_ERR_UNSUPPORTED;
    w.NCH = nch; w.kgroups = kblocks / nch;
    w.d_tiles = (D + 127) / 128;
    w.rpt_b = pe.rpt; w.Gb = pe.G; w.rows_b = pe.rows;
    w.ksteps = (pe.rows + E::UMMA_K - 1) / E::UMMA_K;
    w.total_it = B * pe.G;
    const int base_units = w.d_tiles * w.kgroups;
    int splits = (2 * sm_count() + base_units - 1) / base_units;      // ~2 waves of units
    if (splits > w.total_it) splits = w.total_it;
    if (splits < 1) splits = 1;
    w.it_per_split = (w.total_it + splits - 1) / splits;
    w.splits = (w.total_it + w.it_per_split - 1) / w.it_per_split;
    w.tok_T = dy_tok_T;
    w.dW = dW;
    CUtensorMap tmImg, tmDy;
    if (make_img_tmap(&tmImg, img, E::ESZ, E::DT, B, C, H, W, P, pe.rpt)) return VITK_ERR_TMAP;
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)dy_tok_N, (uint64_t)B};
        uint64_t strides[2] = {(uint64_t)D * E::ESZ, (uint64_t)D * dy_tok_N * E::ESZ};
        uint32_t box[3] = {(uint32_t)(128 / E::ESZ), (uint32_t)pe.rows, 1u};
        if (make_tmap(&tmDy, dy, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, E::DT)) return VITK_ERR_TMAP;
    }
    const int a_bytes = (128 * E::ESZ / 128) * WG_ROWS * 128;
    const int b_bytes = ((w.NCH * WG_ROWS * pe.RB) + 1023) & ~1023;
    const int stage_bytes = a_bytes + b_bytes;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return VITK_ERR_UNSUPPORTED;
    const int smem = stages * stage_bytes + 256 + 1024;
    auto kern = patch_embed_wgrad_kernel<T>;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr_smem = smem;
    }
    kern<<<base_units * w.splits, WG_THREADS, smem, st>>>(tmDy, tmImg, w, pe, stages);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_patch_embed_fwd(const void* img, int img_is_bf16, const void* weight, const float* bias,
                                    const float* pos, long long ldpos, float* out, int B, int C, int H, int W, int P,
                                    int D, int tok_N, int tok_T, void* stream) {
    if (!img || !weight || !pos || !out || ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(weight)) & 15))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (img_is_bf16)
        return pe_fwd_launch<__nv_bfloat16>(img, weight, bias, pos, ldpos, out, B, C, H, W, P, D, tok_N, tok_T, st);
    return pe_fwd_launch<float>(img, weight, bias, pos, ldpos, out, B, C, H, W, P, D, tok_N, tok_T, st);
}

extern "C" int vitk_patch_embed_wgrad(const void* img, int img_is_bf16, const void* dy, int dy_tok_N, int dy_tok_T,
                                      float* dW, int B, int C, int H, int W, int P, int D, void* stream) {
    if (!img || !dy || !dW || ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(dy)) & 15))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // tcgen05 reads MN-major tf32 operands only in the 128B_BASE32B layout, which needs 128-byte patch rows: the weight
    // gradient runs in bf16 (the caller casts an fp32 image / gradient once; still no patch matrix)
    if (!img_is_bf16) return VITK_ERR_UNSUPPORTED;
    return pe_wgrad_launch<__nv_bfloat16>(img, dy, dy_tok_N, dy_tok_T, dW, B, C, H, W, P, D, st);
}

extern "C" int vitk_normalize_u8(const void* x_u8, void* y_bf16, const float* mean, const float* stdv, int B, int C,
                                 int H, int W, void* stream) {
    const long long hw = (long long)H * W, total = (long long)B * C * hw;
    if (!x_u8 || !y_bf16 || !mean || !stdv || total <= 0 || (hw % 16) != 0 ||
        ((reinterpret_cast<uintptr_t>(x_u8) | reinterpret_cast<uintptr_t>(y_bf16)) & 15))
        return VITK_ERR_ARG;
    long long blocks = (total / 16 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    normalize_u8_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint8_t*>(x_u8), reinterpret_cast<__nv_bfloat16*>(y_bf16), mean, stdv, total / 16, C, hw);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
