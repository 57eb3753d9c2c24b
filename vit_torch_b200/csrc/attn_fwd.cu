// Fused flash-style attention forward on tcgen05:  O = softmax(Q K^T * scale) V   per (batch, head).
//
// Replaces on the reference path: timm/DINO Attention.forward core (in-repo witness
// models/swin.py:119-144): attn = (q @ k^T) * scale; softmax; attn @ v -- without materialising [B,H,N,N].
//
// Layout: qkv is the raw output of the qkv Linear, bf16 [B, N, 3, H, d] (= rows [B*N, 3*H*d]); Q/K/V tiles are
// fetched straight out of it with one 2-D TMA map over [B*N, 3*H*d] (box {64, 128}, 128B swizzle; rows >= N of a tile
// belong to the next image and are masked, rows past the tensor end are zero-filled by TMA). O is written bf16 [B, N, H, d] (heads merged = the layout proj consumes),
// lse2[b,h,n] = log2-domain logsumexp of the scaled scores (saved for backward).
//
// One CTA = one (q-block of 128 rows, head, batch). Warps 0-3: softmax (thread == query row, TMEM lane == row),
// warp 4: TMA producer, warp 5: MMA issuer + TMEM allocator. S = Q K_j^T lives in TMEM (128 fp32 columns); P_j is
// written as bf16 into a 128B-swizzled K-major smem tile and multiplied with V_j (MN-major B operand) into a 64-column
// TMEM scratch; the running O is rescaled and accumulated in registers (online softmax).
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int AF_BQ = 128;
constexpr int AF_BKV = 128;
constexpr int AF_THREADS = 192;
constexpr int AF_TILE_BYTES = 128 * 128;                    // one [128 rows x 64 bf16] swizzled tile
constexpr int AF_SMEM_Q = 0;
constexpr int AF_SMEM_K = AF_SMEM_Q + AF_TILE_BYTES;        // 2 stages
constexpr int AF_SMEM_V = AF_SMEM_K + 2 * AF_TILE_BYTES;    // 2 stages
constexpr int AF_SMEM_P = AF_SMEM_V + 2 * AF_TILE_BYTES;    // 2 k-chunks of 64 columns
constexpr int AF_SMEM_BAR = AF_SMEM_P + 2 * AF_TILE_BYTES;  // 114688
constexpr int AF_SMEM_BYTES = AF_SMEM_BAR + 256;
constexpr uint32_t AF_TMEM_COLS = 256;                      // S: [0,128)  O scratch: [128,192)

struct AttnFwdArgs {
    int B, H, N, D;  // D = H * d
    float scale_log2;  // scale * log2(e)
    __nv_bfloat16* out;  // [B*N, D]
    float* lse2;         // [B, H, N]
};

template <int HD>
__global__ void __launch_bounds__(AF_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* v_full = bars + 3;    // [2]
    uint64_t* kv_empty = bars + 5;  // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_full = bars + 9;
    uint64_t* o_full = bars + 10;
    uint64_t* o_free = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qblk * AF_BQ;
    const int nkv = (a.N + AF_BKV - 1) / AF_BKV;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1 KB alignment
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        mbar_init(o_free, 128);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc<AF_TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, AF_TILE_BYTES);
            tma_load_2d(smem + AF_SMEM_Q, &tmQKV, q_full, h * HD, b * a.N + q0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], AF_TILE_BYTES);
                tma_load_2d(smem + AF_SMEM_K + s * AF_TILE_BYTES, &tmQKV, &k_full[s], (a.H + h) * HD, b * a.N + j * AF_BKV);
                mbar_expect_tx(&v_full[s], AF_TILE_BYTES);
                tma_load_2d(smem + AF_SMEM_V + s * AF_TILE_BYTES, &tmQKV, &v_full[s], (2 * a.H + h) * HD, b * a.N + j * AF_BKV);
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
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + AF_SMEM_K + s * AF_TILE_BYTES), 0, 1024);
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
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[s], (j >> 1) & 1);
                if (j > 0) mbar_wait(o_free, (j - 1) & 1);
                tc_fence_after_sync();
                // O_scratch[128, HD] = P[128, kv] * V[kv, HD]: A = P (K-major), B = V tile read MN-major
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);
                const uint32_t v_addr = smem_u32(smem + AF_SMEM_V + s * AF_TILE_BYTES);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t adesc = make_smem_desc_sw128(p_addr + (k >> 2) * AF_TILE_BYTES + (k & 3) * 32, 0, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(v_addr + k * 2048, AF_BKV * 128, 1024);
                    umma_bf16(tmem_o, adesc, bdesc, idesc_pv, k > 0);
                }
                umma_commit(o_full);
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        // ===================== softmax / accumulate (thread == query row) =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        float o_acc[HD];
#pragma unroll
        for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
        uint8_t* p_row = smem + AF_SMEM_P + row * 128;
        const int sw = row & 7;

        for (int j = 0; j < nkv; ++j) {
            const int valid = min(AF_BKV, a.N - j * AF_BKV);
            const int nchunks = (valid + 31) >> 5;
            mbar_wait(s_full, j & 1);
            tc_fence_after_sync();
            // pass 1: row max (masking only in a partially valid chunk; the branch is warp-uniform)
            float mx = -INFINITY;
            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, r);
                tmem_ld_wait();
                if (c * 32 + 32 <= valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        mx = fmaxf(mx, (c * 32 + i < valid) ? __uint_as_float(r[i]) : -INFINITY);
                }
            }
            const float m_new = fmaxf(m_run, mx * a.scale_log2);
            const float alpha = ex2_approx(m_run - m_new);
            float psum = 0.f;
            // pass 2: p = exp2(s*scale_log2 - m_new) -> bf16 -> swizzled smem (A operand of the PV MMA). The loop is
            // issue-bound: one FFMA + one MUFU.EX2 + one FADD per element, one cvt per pair.
            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, r);
                tmem_ld_wait();
                float p[32];
                if (c * 32 + 32 <= valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        p[i] = ex2_approx(fmaf(__uint_as_float(r[i]), a.scale_log2, -m_new));
                        psum += p[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = ex2_approx(fmaf(__uint_as_float(r[i]), a.scale_log2, -m_new));
                        p[i] = (c * 32 + i < valid) ? v : 0.f;
                        psum += p[i];
                    }
                }
                uint8_t* dst = p_row + (c >> 1) * AF_TILE_BYTES;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint4 pk = make_uint4(pack_bf16(p[u * 8 + 0], p[u * 8 + 1]), pack_bf16(p[u * 8 + 2], p[u * 8 + 3]),
                                                pack_bf16(p[u * 8 + 4], p[u * 8 + 5]), pack_bf16(p[u * 8 + 6], p[u * 8 + 7]));
                    const int unit = (c & 1) * 4 + u;
                    *reinterpret_cast<uint4*>(dst + ((unit ^ sw) << 4)) = pk;
                }
            }
            tc_fence_before_sync();
            mbar_arrive(s_free);
            fence_proxy_async_smem();
            mbar_arrive(p_full);
            l_run = l_run * alpha + psum;
            m_run = m_new;
            // accumulate O
            mbar_wait(o_full, j & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_o + lane_off + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] = fmaf(o_acc[c * 16 + i], alpha, __uint_as_float(r[i]));
            }
            tc_fence_before_sync();
            mbar_arrive(o_free);
        }
        const int n = q0 + row;
        if (n < a.N) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = a.out + ((long long)b * a.N + n) * a.D + h * HD;
#pragma unroll
            for (int u = 0; u < HD / 8; ++u) {
                const uint4 pk = make_uint4(pack_bf16(o_acc[u * 8 + 0] * inv, o_acc[u * 8 + 1] * inv),
                                            pack_bf16(o_acc[u * 8 + 2] * inv, o_acc[u * 8 + 3] * inv),
                                            pack_bf16(o_acc[u * 8 + 4] * inv, o_acc[u * 8 + 5] * inv),
                                            pack_bf16(o_acc[u * 8 + 6] * inv, o_acc[u * 8 + 7] * inv));
                st_v4(dst + u * 8, pk);
            }
            a.lse2[((long long)b * a.H + h) * a.N + n] = m_run + log2f(l_run);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc<AF_TMEM_COLS>(tmem_base);
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

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_attn_fwd(const void* qkv_bf16, void* out_bf16, float* lse2, int B, int N, int H, int d,
                             float scale, void* stream) {
    if (B <= 0 || N <= 0 || H <= 0 || !(d == 64 || d == 48) || !qkv_bf16 || !out_bf16 || !lse2) return VITK_ERR_ARG;
    CUtensorMap tm;
    if (make_tok_tmap2d(&tm, qkv_bf16, (long long)B * N, 3LL * H * d, 128)) return VITK_ERR_TMAP;
    AttnFwdArgs a;
    a.B = B; a.H = H; a.N = N; a.D = H * d;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    a.lse2 = lse2;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((N + AF_BQ - 1) / AF_BQ, H, B);
    static bool attr64 = false, attr48 = false;
    if (d == 64) {
        if (!attr64) {
            if (cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES) != cudaSuccess)
                return VITK_ERR_CUDA;
            attr64 = true;
        }
        attn_fwd_kernel<64><<<grid, AF_THREADS, AF_SMEM_BYTES, st>>>(tm, a);
    } else {
        if (!attr48) {
            if (cudaFuncSetAttribute(attn_fwd_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES) != cudaSuccess)
                return VITK_ERR_CUDA;
            attr48 = true;
        }
        attn_fwd_kernel<48><<<grid, AF_THREADS, AF_SMEM_BYTES, st>>>(tm, a);
    }
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
