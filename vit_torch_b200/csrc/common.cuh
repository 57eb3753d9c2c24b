// Shared sm_100a device helpers: mbarrier, TMA, tcgen05 (UMMA/TMEM) PTX wrappers,
// shared-memory matrix descriptors, 128-bit vector access and warp reductions.
// Everything here is inline PTX for Blackwell (sm_100a); there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vitk {

// ----------------------------------------------------------------------------------------------
// basic helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ----------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): one issue slot per two elements
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t f2_bcast(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ----------------------------------------------------------------------------------------------
// exact-erf GELU (nn.GELU default) and its derivative, fp32.
//   gelu(x) = x * Phi(x),  gelu'(x) = Phi(x) + x * phi(x)
// With Q(|x|) = 1 - Phi(|x|) = exp(-x^2/2) * P(|x|), P = Mills ratio / sqrt(2 pi):
//   gelu(x)  = max(x, 0) - |x| * Q
//   gelu'(x) = step(x) - sign(x) * exp(-x^2/2) * (P(|x|) - |x| / sqrt(2 pi))
// P is a degree-8 polynomial on [0, 6] (weighted minimax fit, |error in Q| and |x| * |error in Q| <= 1.8e-6 -- three
// orders below bf16 resolution of the outputs; beyond 6 the argument is clamped, exp(-18) = 1.5e-8 makes Q vanish).
// One MUFU (ex2) per element; the polynomial runs on packed FFMA2. `gelu_pairs<NP>` evaluates NP pairs in lock step
// (step-major order) so that the dependent Horner chains of different pairs interleave in the instruction stream.
// ----------------------------------------------------------------------------------------------
template <int NP, bool WITH_GRAD>
__device__ __forceinline__ void gelu_pairs(const float* x, float* g, float* d) {
    constexpr float C0 = 4.999982417e-01f, C1 = -3.988357782e-01f, C2 = 2.489434332e-01f, C3 = -1.289402843e-01f,
                    C4 = 5.464975536e-02f, C5 = -1.768272184e-02f, C6 = 3.933028784e-03f, C7 = -5.189787480e-04f,
                    C8 = 3.003616439e-05f;
    constexpr float kInvSqrt2Pi = 0.3989422804014327f;
    uint64_t xx[NP], ax[NP], axc[NP], e[NP], p[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const float a0 = fabsf(x[2 * i]), a1 = fabsf(x[2 * i + 1]);
        xx[i] = f2_pack(x[2 * i], x[2 * i + 1]);
        ax[i] = f2_pack(a0, a1);
        axc[i] = f2_pack(fminf(a0, 6.0f), fminf(a1, 6.0f));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) e[i] = f2_mul(f2_mul(xx[i], xx[i]), f2_bcast(-0.72134752044448170f));  // -x^2/2 log2 e
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float e0, e1;
        f2_unpack(e[i], e0, e1);
        e[i] = f2_pack(ex2_approx_(e0), ex2_approx_(e1));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(f2_bcast(C8), axc[i], f2_bcast(C7));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C6));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C5));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C4));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C3));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C2));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C1));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = f2_fma(p[i], axc[i], f2_bcast(C0));
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const uint64_t q = f2_mul(e[i], p[i]);  // Q(|x|)
        const uint64_t nq = f2_mul(q, f2_bcast(-1.0f));
        f2_unpack(f2_fma(ax[i], nq, f2_pack(fmaxf(x[2 * i], 0.f), fmaxf(x[2 * i + 1], 0.f))), g[2 * i], g[2 * i + 1]);
    }
    if constexpr (WITH_GRAD) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const uint64_t w = f2_mul(e[i], f2_fma(ax[i], f2_bcast(-kInvSqrt2Pi), p[i]));  // Q - |x| phi
            const uint64_t step = f2_pack(x[2 * i] >= 0.f ? 1.f : 0.f, x[2 * i + 1] >= 0.f ? 1.f : 0.f);
            const uint64_t nsgn = f2_fma(step, f2_bcast(-2.0f), f2_bcast(1.0f));  // -sign(x)
            f2_unpack(f2_fma(w, nsgn, step), d[2 * i], d[2 * i + 1]);
        }
    }
}
template <bool WITH_GRAD>
__device__ __forceinline__ void gelu_pair(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
    const float x[2] = {x0, x1};
    float g[2], d[2] = {0.f, 0.f};
    gelu_pairs<1, WITH_GRAD>(x, g, d);
    g0 = g[0]; g1 = g[1]; d0 = d[0]; d1 = d[1];
}
__device__ __forceinline__ float gelu_f(float x) {
    float g0, g1, d0, d1;
    gelu_pair<false>(x, x, g0, g1, d0, d1);
    return g0;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    float g0, g1, d0, d1;
    gelu_pair<true>(x, x, g0, g1, d0, d1);
    return d0;
}

// single-instruction exp2 (MUFU.EX2): the softmax loops are issue-bound, exp2f() costs several instructions more
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// Wait used by the single-thread roles (TMA producer, MMA issuer) of kernels whose elementwise warps are the critical
// path: a bare try_wait loop issues three instructions per poll from a warp that is always eligible and takes issue
// slots from the elementwise warps of the same scheduler (24 % of all warp instructions of attn_bwd_dkv2 were such
// polls); sleeping between polls gives the slots back for at most ~32 ns of extra wake-up latency.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

// The long waits of the GEMM kernels' single-thread roles (producer: a free smem stage; MMA issuer: a free accumulator).
// In the epilogue-bound instantiations (GELU / GELU' / residual) the bare polling loops of these two warps were 16 % of
// all warp instructions and shared their schedulers with two of the eight epilogue warps (ncu, profiles/r02_summary.md);
// VITK_GEMM_DBG bit 16 restores the bare loop for A/B runs.
__device__ __forceinline__ void mbar_wait_role(uint64_t* bar, uint32_t parity, int dbg) {
    if (dbg & 16) mbar_wait(bar, parity);
    else mbar_wait_backoff(bar, parity);
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), loads complete on an mbarrier, coordinates innermost first
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}
// L2 prefetch of a tile (no smem destination, no barrier): turns the later TMA load of the same box into an L2 hit
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// TMA store smem -> global (bulk group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* m, uint32_t src_saddr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src_saddr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols> __device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
    static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM cols: pow2 in [32,512]");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: the two CTAs of a 2-CTA cluster execute one UMMA of M = 256 (each CTA owns 128
// rows of A and D and half of the B tile). Shared-memory addresses with bit 24 cleared name the same offset in the
// leader (even-ranked) CTA of the pair.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t kCols> __device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {
    static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM cols: pow2 in [32,512]");
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols> __device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the leader CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// arrives on `bar` in BOTH CTAs of the pair once all previously issued cta_group::2 MMAs have completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Instruction descriptor for kind::f16 with BF16 A/B and FP32 accumulator.
// bits: [4,6) c_format=1(F32) | [7,10) a_format=1(BF16) | [10,13) b_format=1(BF16) | [15] a_major | [16] b_major |
//       [17,23) N>>3 | [24,29) M>>4          (major: 0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// Shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).
//   bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
// K-major tile  (rows of 64 bf16 = 128 B, TMA box {64, rows}):  SBO = 1024 (8-row group), LBO unused.
//   advance along K by 16 elements = +32 B on the start address.
// MN-major tile (TMA box {64 mn, BK k-rows} per 64-wide MN chunk): SBO = 1024 (8 k-rows), LBO = BK*128 (next MN chunk).
//   advance along K by 16 rows = +2048 B on the start address.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// TMEM -> registers: each thread of the warp reads its own lane (32*(warp%4)+lane), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM (same lane/column addressing as tmem_ld_32x32b_x16)
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// explicit shared-space vector access (32-bit shared addresses): keeps the compiler from falling back to generic
// LD/ST when the address space of a carved-up dynamic smem pointer is not provable
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t saddr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

// ----------------------------------------------------------------------------------------------
// global memory vector access
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Column sums over the 32 rows of a warp for the 16 columns of a chunk (lane = row): recursive halving, 16 shuffles.
// On return lane l (l even) holds the total of column ((l >> 1) & 15) ... see `col_of_lane`.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = up ? v[i] : v[i + 8];
            const float keep = up ? v[i + 8] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float b4[4];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? a[i] : a[i + 4];
            const float keep = up ? a[i + 4] : a[i];
            b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float c2[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? b4[i] : b4[i + 2];
            const float keep = up ? b4[i + 2] : b4[i];
            c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float d;
    {
        const bool up = lane & 2;
        const float send = up ? c2[0] : c2[1];
        const float keep = up ? c2[1] : c2[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;  // column index held by this lane: 8*bit4 + 4*bit3 + 2*bit2 + bit1 of `lane`
}
__device__ __forceinline__ int warp_colsum16_col(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}


// ---- optional per-CTA event trace (instrumented build only: -DVITK_TRACE, scripts/trace_attn.py) ----
// trace[cta * 64 + id] = SM clock at event `id`; all roles of a CTA share one SM clock, so differences are cycles.
#ifdef VITK_TRACE
__device__ __forceinline__ void trace_ev(long long* trace, int id) {
    if (trace != nullptr) {
        const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (cta < 16384u) trace[cta * 64 + id] = clock64();
    }
}
#define VITK_TRACE_EV(ptr, id) trace_ev((ptr), (id))
#else
#define VITK_TRACE_EV(ptr, id) ((void)0)
#endif

}  // namespace vitk
