// Host launcher + C-ABI entry for the tcgen05 GEMM (see gemm.cuh for the kernel).
#include "gemm.cuh"
#include "gemm2.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"
#include <stdlib.h>

namespace vitk {

// Choose the split-K factor for accumulate-epilogue GEMMs (wgrad): minimise waves * k-blocks-per-unit, with a fixed
// per-unit overhead (pipeline fill + epilogue) expressed in k-block equivalents.
static int choose_splits(int tiles, int num_kblocks, int sms) {
    int best = 1;
    long long best_cost = -1;
    const int max_s = num_kblocks < 64 ? num_kblocks : 64;
    for (int s = 1; s <= max_s; ++s) {
        const int kbps = (num_kblocks + s - 1) / s;
        const int s_eff = (num_kblocks + kbps - 1) / kbps;
        const long long units = (long long)tiles * s_eff;
        const long long waves = (units + sms - 1) / sms;
        const long long cost = waves * (kbps + 10);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = s_eff;
        }
    }
    return best;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                       const GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, EPI>;
    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr_set = true;
    }
    const int units = g.num_m_tiles * g.num_n_tiles * g.splits * g.nbatch_h * g.nbatch_b;
    const int grid = units < sm_count() ? units : sm_count();
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmO, tmO2, g);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

// ---- CTA-pair kernel (gemm2.cuh) ----
static int max_clusters() { return sm_count() / 2 > 0 ? sm_count() / 2 : 1; }

// VITK_GEMM_2CTA=0 keeps every GEMM on the 1-CTA kernel (A/B measurements)
static bool use_2cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VITK_GEMM_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                        const GemmArgs& g, cudaStream_t stream) {
    using Cfg = Gemm2Cfg<EPI>;
    auto kern = gemm2_bf16_kernel<A_MN, B_MN, EPI>;
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr_set = true;
    }
    const int units = ((g.num_m_tiles + 1) / 2) * g.num_n_tiles * g.splits;
    const int clusters = units < max_clusters() ? units : max_clusters();
    kern<<<2 * clusters, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmO, tmO2, g);  // static __cluster_dims__(2,1,1)
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

static int dispatch_gemm2(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const CUtensorMap& tmO, const CUtensorMap& tmO2, const GemmArgs& g, cudaStream_t s) {
#define VITK_CASE2(AM, BM_, E)                       \
    if (a_mn == (AM) && b_mn == (BM_) && epi == (E)) \
        return launch_gemm2<(AM) != 0, (BM_) != 0, (E)>(tmA, tmB, tmO, tmO2, g, s);
    VITK_CASE2(0, 0, EPI_STORE_BF16)
    VITK_CASE2(0, 0, EPI_BIAS_GELU)
    VITK_CASE2(0, 0, EPI_RESID_F32)
    VITK_CASE2(0, 1, EPI_STORE_BF16)
    VITK_CASE2(0, 1, EPI_DGELU)
    VITK_CASE2(1, 1, EPI_ATOMIC_F32)
#undef VITK_CASE2
    return VITK_ERR_UNSUPPORTED;
}
static bool gemm2_supported(int a_mn, int b_mn, int epi) {
    return (a_mn == 0 && b_mn == 0 && (epi == EPI_STORE_BF16 || epi == EPI_BIAS_GELU || epi == EPI_RESID_F32)) ||
           (a_mn == 0 && b_mn == 1 && (epi == EPI_STORE_BF16 || epi == EPI_DGELU)) ||
           (a_mn == 1 && b_mn == 1 && epi == EPI_ATOMIC_F32);
}

template <int BN>
static int dispatch_gemm(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const CUtensorMap& tmO, const CUtensorMap& tmO2, const GemmArgs& g, cudaStream_t s) {
#define VITK_CASE(AM, BM_, E)                        \
    if (a_mn == (AM) && b_mn == (BM_) && epi == (E)) \
        return launch_gemm<BN, (AM) != 0, (BM_) != 0, (E)>(tmA, tmB, tmO, tmO2, g, s);
    // forward Linear: activations K-major, weight [N,K] K-major
    VITK_CASE(0, 0, EPI_STORE_BF16)
    VITK_CASE(0, 0, EPI_BIAS_GELU)
    VITK_CASE(0, 0, EPI_RESID_F32)
    VITK_CASE(0, 0, EPI_STORE_F32)
    VITK_CASE(0, 0, EPI_TOKENS_F32)
    // dgrad: dY K-major, weight [N_out, K_in] read as MN-major B
    VITK_CASE(0, 1, EPI_STORE_BF16)
    VITK_CASE(0, 1, EPI_DGELU)
    VITK_CASE(0, 1, EPI_STORE_F32)
    VITK_CASE(0, 1, EPI_RESID_F32)
    // wgrad: dY^T and X^T, both MN-major, split-K accumulate
    VITK_CASE(1, 1, EPI_ATOMIC_F32)
    VITK_CASE(1, 1, EPI_STORE_F32)
    VITK_CASE(1, 1, EPI_STORE_BF16)
    VITK_CASE(1, 0, EPI_STORE_F32)
#undef VITK_CASE
    return VITK_ERR_UNSUPPORTED;
}

}  // namespace vitk

using namespace vitk;

struct GemmBatch {
    int nh = 1, nb = 1;
    long long sa_h = 0, sa_b = 0, sb_h = 0, sb_b = 0, so_h = 0, so_b = 0;  // element strides
};

static int gemm_impl(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, int M,
                     int N, int K, int epilogue, const float* bias, const float* gamma, const float* resid,
                     long long ldr, void* out, long long ldo, void* out2, long long ldo2, const void* aux,
                     long long ldaux, int splits, const float* rowscale, int rows_per_sample, int tok_n, int tok_N,
                     int tok_T, void* stream, const GemmBatch& bt = GemmBatch(), float* colsum = nullptr) {
    if (M <= 0 || N <= 0 || K <= 0) return VITK_ERR_ARG;
    if ((lda % 8) != 0 || (ldb % 8) != 0) return VITK_ERR_ARG;
    // The epilogue works on 4-column groups. N that is not a multiple of 4 is allowed for the plain store epilogues when
    // the output pitch has room for the rounded-up width: the extra columns come out as zeros (TMA zero-fills B rows >= N).
    const int Nepi = (N + 3) & ~3;
    if (Nepi != N && (ldo < Nepi || bias || gamma || resid || aux || out2 ||
                      !(epilogue == EPI_STORE_BF16 || epilogue == EPI_STORE_F32)))
        return VITK_ERR_ARG;
    if ((epilogue == EPI_STORE_BF16 || epilogue == EPI_BIAS_GELU || epilogue == EPI_DGELU) && (ldo % 4) != 0)
        return VITK_ERR_ARG;
    if (A == nullptr || B == nullptr) return VITK_ERR_ARG;
    if (out == nullptr && !(epilogue == EPI_BIAS_GELU && out2 != nullptr)) return VITK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out)) & 15)
        return VITK_ERR_ARG;
    if (rowscale != nullptr && rows_per_sample <= 0) return VITK_ERR_ARG;
    if (epilogue == EPI_TOKENS_F32 && (tok_n <= 0 || tok_N < tok_n + tok_T || resid == nullptr)) return VITK_ERR_ARG;
    const int BN = (Nepi % 256 == 0 || Nepi >= 1024) ? 256 : 128;
    // large plain GEMMs run on the CTA-pair kernel (256 x 256 tiles per 2-CTA cluster)
    // ... when wave quantisation does not eat the gain: a pair unit is 256 rows, so e.g. 197 row tiles x 3 column tiles
    // are 297 units on 74 clusters (4.01 waves) against 591 tiles on 148 SMs (3.99 waves). Measured (sustained clocks):
    // +7 % on N = 2304 / 3072 forward GEMMs, -6 % on N = 768, wgrad (MN-major A, split-K) even.
    bool pair = use_2cta() && BN == 256 && bt.nh * bt.nb == 1 && M >= 1024 && a_mn_major == 0 &&
                gemm2_supported(a_mn_major != 0, b_mn_major != 0, epilogue);
    if (pair) {
        const long long mt = (M + GEMM_BM - 1) / GEMM_BM, nt = (Nepi + BN - 1) / BN;
        const long long u1 = mt * nt, u2 = ((mt + 1) / 2) * nt;
        const long long w1 = (u1 + sm_count() - 1) / sm_count(), w2 = (u2 + max_clusters() - 1) / max_clusters();
        const double eff1 = (double)u1 / (double)(w1 * sm_count()), eff2 = (double)(mt * nt) / (double)(w2 * 2 * max_clusters());
        pair = eff2 >= 0.9 * eff1;
    }

    GemmArgs g;
    g.M = M; g.N = Nepi; g.K = K;
    g.num_m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    g.num_n_tiles = (Nepi + BN - 1) / BN;
    g.num_kblocks = (K + GEMM_BK - 1) / GEMM_BK;
    const bool accumulate = (epilogue == EPI_ATOMIC_F32);
    int s = 1;
    if (accumulate)
        s = splits > 0 ? splits
                       : (pair ? choose_splits(((g.num_m_tiles + 1) / 2) * g.num_n_tiles, g.num_kblocks, max_clusters())
                               : choose_splits(g.num_m_tiles * g.num_n_tiles, g.num_kblocks, sm_count()));
    if (s > g.num_kblocks) s = g.num_kblocks;
    g.kblocks_per_split = (g.num_kblocks + s - 1) / s;
    g.splits = (g.num_kblocks + g.kblocks_per_split - 1) / g.kblocks_per_split;
    g.bias = bias; g.gamma = gamma; g.resid = resid; g.ldr = ldr;
    g.out = out; g.ldo = ldo; g.out2 = out2; g.ldo2 = ldo2;
    g.aux = reinterpret_cast<const __nv_bfloat16*>(aux); g.ldaux = ldaux;
    g.rowscale = rowscale; g.rows_per_sample = rows_per_sample;
    g.tok_n = tok_n; g.tok_N = tok_N; g.tok_T = tok_T;
    g.nbatch_h = bt.nh; g.nbatch_b = bt.nb; g.so_h = bt.so_h; g.so_b = bt.so_b;
    g.colsum = colsum;
    g.pe_G = 0; g.pe_rows = 0;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("VITK_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
        g.dbg = dbg;
    }
    if (colsum != nullptr && !(epilogue == EPI_STORE_BF16 || epilogue == EPI_DGELU)) return VITK_ERR_UNSUPPORTED;
    if (bt.nh < 1 || bt.nb < 1) return VITK_ERR_ARG;
    if ((bt.sa_h | bt.sa_b | bt.sb_h | bt.sb_b) % 8 != 0) return VITK_ERR_ARG;
    if (bt.nh * bt.nb > 1 && (accumulate || !(epilogue == EPI_STORE_BF16 || epilogue == EPI_STORE_F32) || bias != nullptr))
        return VITK_ERR_UNSUPPORTED;
    if (epilogue == EPI_BIAS_GELU && out2 == nullptr) return VITK_ERR_ARG;
    if (epilogue == EPI_DGELU && aux == nullptr) return VITK_ERR_ARG;

    CUtensorMap tmA, tmB;
    int rc;
    // K-major operand: global [rows, K], box {64 k, rows};  MN-major operand: global [K, rows], box {64 rows, 64 k}
    const uint64_t nh = bt.nh, nb = bt.nb;
    if (nh * nb == 1) {
        g.a_perm[0] = g.b_perm[0] = 0; g.a_perm[1] = g.b_perm[1] = 1; g.a_perm[2] = g.b_perm[2] = 2;
        if (!a_mn_major) rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM);
        else             rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BK);
        if (rc) return VITK_ERR_TMAP;
        if (!b_mn_major) rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK,
                                                (uint32_t)(pair ? BN / 2 : BN));  // pair: each CTA loads half of the B tile
        else             rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BK);
        if (rc) return VITK_ERR_TMAP;
    } else {
    if (!a_mn_major) rc = make_tmap_4d_bf16(&tmA, g.a_perm, A, K, M, nh, nb, lda, bt.sa_h, bt.sa_b, GEMM_BK, GEMM_BM);
    else             rc = make_tmap_4d_bf16(&tmA, g.a_perm, A, M, K, nh, nb, lda, bt.sa_h, bt.sa_b, 64, GEMM_BK);
    if (rc) return VITK_ERR_TMAP;
    if (!b_mn_major) rc = make_tmap_4d_bf16(&tmB, g.b_perm, B, K, N, nh, nb, ldb, bt.sb_h, bt.sb_b, GEMM_BK, (uint32_t)BN);
    else             rc = make_tmap_4d_bf16(&tmB, g.b_perm, B, N, K, nh, nb, ldb, bt.sb_h, bt.sb_b, 64, GEMM_BK);
    if (rc) return VITK_ERR_TMAP;
    }

    // outputs through TMA stores (non-batched, 16-byte aligned rows): [32 rows x 64 B] boxes per epilogue warp
    CUtensorMap tmO, tmO2;
    memset(&tmO, 0, sizeof(tmO));
    memset(&tmO2, 0, sizeof(tmO2));
    g.tma_out = 0;
    {
        static int tma_off = -1;
        if (tma_off < 0) { const char* e = getenv("VITK_GEMM_TMA_STORE"); tma_off = (e != nullptr && e[0] == '0') ? 1 : 0; }
        const bool f32out = (epilogue == EPI_RESID_F32 || epilogue == EPI_STORE_F32);
        const int esz = f32out ? 4 : 2;
        auto ok16 = [](const void* p_, long long ld, int es) {
            return p_ != nullptr && ((reinterpret_cast<uintptr_t>(p_) | (uintptr_t)(ld * es)) & 15) == 0;
        };
        bool use = !tma_off && nh * nb == 1 &&
                   (epilogue == EPI_STORE_BF16 || epilogue == EPI_BIAS_GELU || epilogue == EPI_RESID_F32 ||
                    epilogue == EPI_DGELU || epilogue == EPI_STORE_F32);
        if (use) {
            // BIAS_GELU: `out` (gelu') is optional, `out2` (gelu) is the mandatory bf16 output
            if (epilogue == EPI_BIAS_GELU) use = ok16(out2, ldo2, 2) && (out == nullptr || ok16(out, ldo, 2));
            else use = ok16(out, ldo, esz);
        }
        if (use) {
            if (epilogue == EPI_BIAS_GELU) {
                if (out != nullptr && make_tmap_2d_store(&tmO, out, 2, (uint64_t)Nepi, (uint64_t)M, (uint64_t)ldo, 32))
                    return VITK_ERR_TMAP;
                if (make_tmap_2d_store(&tmO2, out2, 2, (uint64_t)Nepi, (uint64_t)M, (uint64_t)ldo2, 32)) return VITK_ERR_TMAP;
            } else {
                if (make_tmap_2d_store(&tmO, out, esz, (uint64_t)Nepi, (uint64_t)M, (uint64_t)ldo, 32)) return VITK_ERR_TMAP;
            }
            g.tma_out = 1;
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (pair) return dispatch_gemm2(a_mn_major != 0, b_mn_major != 0, epilogue, tmA, tmB, tmO, tmO2, g, st);
    if (BN == 256) return dispatch_gemm<256>(a_mn_major != 0, b_mn_major != 0, epilogue, tmA, tmB, tmO, tmO2, g, st);
    return dispatch_gemm<128>(a_mn_major != 0, b_mn_major != 0, epilogue, tmA, tmB, tmO, tmO2, g, st);
}

extern "C" int vitk_set_sm_limit(int n) {
    if (n < 0) return VITK_ERR_ARG;
    sm_limit_ref() = n;
    return VITK_OK;
}

extern "C" int vitk_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                              int b_mn_major, int M, int N, int K, int epilogue, const float* bias, const float* gamma,
                              const float* resid, long long ldr, void* out, long long ldo, void* out2, long long ldo2,
                              const void* aux, long long ldaux, int splits, void* stream) {
    return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epilogue, bias, gamma, resid, ldr, out, ldo, out2,
                     ldo2, aux, ldaux, splits, nullptr, 0, 0, 0, 0, stream);
}

extern "C" int vitk_gemm_bf16_ex(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                                 int b_mn_major, int M, int N, int K, int epilogue, const float* bias,
                                 const float* gamma, const float* resid, long long ldr, void* out, long long ldo,
                                 void* out2, long long ldo2, const void* aux, long long ldaux, int splits,
                                 const float* rowscale, int rows_per_sample, int tok_n, int tok_N, int tok_T,
                                 float* colsum, void* stream) {
    return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epilogue, bias, gamma, resid, ldr, out, ldo, out2,
                     ldo2, aux, ldaux, splits, rowscale, rows_per_sample, tok_n, tok_N, tok_T, stream, GemmBatch(),
                     colsum);
}

extern "C" int vitk_gemm_bf16_batched(const void* A, long long lda, long long sa_h, long long sa_b, int a_mn_major,
                                      const void* B, long long ldb, long long sb_h, long long sb_b, int b_mn_major,
                                      int M, int N, int K, int nbatch_h, int nbatch_b, int epilogue, void* out,
                                      long long ldo, long long so_h, long long so_b, void* stream) {
    GemmBatch bt;
    bt.nh = nbatch_h; bt.nb = nbatch_b;
    bt.sa_h = sa_h; bt.sa_b = sa_b; bt.sb_h = sb_h; bt.sb_b = sb_b; bt.so_h = so_h; bt.so_b = so_b;
    return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epilogue, nullptr, nullptr, nullptr, 0, out, ldo,
                     nullptr, 0, nullptr, 0, 1, nullptr, 0, 0, 0, 0, stream, bt);
}
