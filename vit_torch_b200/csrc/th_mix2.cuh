// Talking-heads mixing, version 2 (models/cait.py:116-125): one WARP per (image, query row), no block barriers.
//
//   forward : S' = scale*Wl S + bl (over the head axis) ; P = softmax_j(S') ; P' = Ww P + bw
//   backward: dP = Ww^T dP' ; dS' = P o (dP - rowsum(dP o P)) ; dS = scale * Wl^T dS'
//             dWw = sum dP' P^T ; dbw = sum dP' ; dWl = scale * sum dS' S^T ; dbl = sum dS'
//
// Every head mix is a [16 columns x H] x [H x H] product and runs on the tensor cores through mma.sync.m16n8k8 (tf32):
// the 16 x H tile of logits / probabilities lives in the MMA fragment layout (lane = (gid, tig): columns 2gid, 2gid+1 of
// the tile, heads tig / tig+4 in A layout, 2tig / 2tig+1 in C layout), loads and stores are 8-byte / 4-byte accesses
// that cover whole 32-byte sectors, and chained mixes reuse the accumulator registers as the next A operand by
// permuting the K slots (slot tig <-> head 2tig, slot tig+4 <-> head 2tig+1) -- no shuffles. The logit mix is done in
// 3xTF32 (hi/lo split of S and Wl: fp32-level accuracy before the exponential), the probability / gradient mixes in
// plain tf32. The weight gradients (2 H^2 dot products over all (b, i, j)) are [H x 16 columns] x [16 columns x H]
// products of the transposed tiles, staged through a per-warp shared-memory scratch and accumulated in registers
// over all rows a warp handles; one atomic per entry per thread block at the end.
// Row statistics use shuffles over the 8 lanes that share a head. Memory-bound by design: forward reads S (fp32) and
// writes P' (bf16); backward reads S (fp32), dP' (bf16) and writes dS (bf16).
#pragma once
#include "common.cuh"

namespace vitk {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo with hi, lo representable in tf32 (3xTF32 operand split)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

constexpr int TH2_WARPS = 8;
constexpr int TH2_MAX_TILES = 13;   // rows of up to 208 keys (CaiT at 224 px: 196)

template <int H> struct Th2 {
    static constexpr int KS = (H + 7) / 8;   // 8-head groups: K steps of a mix / N tiles of its output
    static constexpr int HP = KS * 8;
};

// B fragments of a head-mixing matrix: b[ks][nt] for D[col][out] += A[col][in] * M[in][out], `in` addressed by K slot.
//   slot_perm = false: K slot s of step ks is input head 8ks + s            (A operand loaded from memory)
//   slot_perm = true : K slot tig is head 8ks + 2tig, slot tig+4 head 8ks + 2tig + 1 (A operand = a C fragment)
// get(in, out) returns M[in][out] (0 outside H).
template <int H, bool SLOT_PERM, class F>
__device__ __forceinline__ void th2_bfrag(uint32_t (&b)[Th2<H>::KS][Th2<H>::KS][2], int gid, int tig, F get) {
#pragma unroll
    for (int ks = 0; ks < Th2<H>::KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < Th2<H>::KS; ++nt) {
            const int in0 = 8 * ks + (SLOT_PERM ? 2 * tig : tig), in1 = 8 * ks + (SLOT_PERM ? 2 * tig + 1 : tig + 4);
            const int out = 8 * nt + gid;
            b[ks][nt][0] = __float_as_uint((in0 < H && out < H) ? get(in0, out) : 0.f);
            b[ks][nt][1] = __float_as_uint((in1 < H && out < H) ? get(in1, out) : 0.f);
        }
}

// C fragment (heads 2tig, 2tig+1 of group nt; columns 2gid, 2gid+1) -> A fragment of the next mix (slot-permuted K)
__device__ __forceinline__ void th2_c_to_a(const float (&c)[4], uint32_t (&a)[4]) {
    a[0] = __float_as_uint(c[0]);  // (col 2gid,   head 2tig)   -> (row gid,   slot tig)
    a[1] = __float_as_uint(c[2]);  // (col 2gid+1, head 2tig)   -> (row gid+8, slot tig)
    a[2] = __float_as_uint(c[1]);  // (col 2gid,   head 2tig+1) -> (row gid,   slot tig+4)
    a[3] = __float_as_uint(c[3]);  // (col 2gid+1, head 2tig+1) -> (row gid+8, slot tig+4)
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
// logits of two adjacent key columns of one head: fp32 plane (8-byte load) or bf16 plane (4-byte load; bf16 values are
// exact in tf32, so the 3xTF32 logit mix needs no low part of S)
template <typename ST> __device__ __forceinline__ float2 th2_load_s2(const ST* p) {
    if constexpr (sizeof(ST) == 4) {
        return __ldg(reinterpret_cast<const float2*>(p));
    } else {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
        return make_float2(bf16_lo(v), bf16_hi(v));
    }
}

// the same pair kept in the registers it was loaded into (bf16 planes: one packed 32-bit register)
template <typename ST> struct Th2Raw { using type = float2; };
template <> struct Th2Raw<__nv_bfloat16> { using type = uint32_t; };
template <typename ST> __device__ __forceinline__ typename Th2Raw<ST>::type th2_load_raw(const ST* p) {
    if constexpr (sizeof(ST) == 4) return __ldg(reinterpret_cast<const float2*>(p));
    else return __ldg(reinterpret_cast<const uint32_t*>(p));
}
__device__ __forceinline__ float2 th2_raw_f2(float2 v) { return v; }
__device__ __forceinline__ float2 th2_raw_f2(uint32_t v) { return make_float2(bf16_lo(v), bf16_hi(v)); }

// Per-lane element offsets inside a (b, i) row group of the planes: heads in A layout (tig, tig+4: what the lane loads)
// and in C layout (2tig, 2tig+1: what it stores), first column 2*gid of a tile. Tile t adds the constant 16*t, so the
// unrolled row needs no address arithmetic per access (it was 18 % of the forward kernel's instructions).
template <int H> struct Th2Off {
    static constexpr int KS = Th2<H>::KS;
    long long a[KS][2], c[KS][2];
    bool a_ok[KS][2], c_ok[KS][2];
    __device__ __forceinline__ Th2Off(long long plane, int gid, int tig) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ha = 8 * ks + tig + 4 * e, hc = 8 * ks + 2 * tig + e;
                a_ok[ks][e] = (H % 8 == 0) || ha < H;
                c_ok[ks][e] = (H % 8 == 0) || hc < H;
                a[ks][e] = (a_ok[ks][e] ? ha : 0) * plane + 2 * gid;
                c[ks][e] = (c_ok[ks][e] ? hc : 0) * plane + 2 * gid;
            }
    }
};

struct Th2True { static constexpr bool value = true; };
struct Th2False { static constexpr bool value = false; };

template <int H, int NT, typename ST>
__global__ void __launch_bounds__(TH2_WARPS * 32)
th_mix2_fwd_kernel(const ST* __restrict__ S, const float* __restrict__ wl, const float* __restrict__ bl,
                   const float* __restrict__ ww, const float* __restrict__ bw, float scale,
                   __nv_bfloat16* __restrict__ Pm, float* __restrict__ rowmax, float* __restrict__ rowsum, int B, int N,
                   int Np) {
    constexpr int KS = Th2<H>::KS;
    using RawS = typename Th2Raw<ST>::type;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long plane = (long long)N * Np;
    constexpr float kLog2e = 1.4426950408889634f;
    const Th2Off<H> off(plane, gid, tig);
    const int nfull = N >> 4;    // tiles whose 16 columns are all real keys: no masks, no bounds checks (N <= Np)

    // mixing matrices as B fragments (constant over the kernel). Logit mix in the log2 domain: scale*log2e folded in.
    uint32_t b1h[KS][KS][2], b1l[KS][KS][2], b2[KS][KS][2];
    {
        uint32_t t[KS][KS][2];
        th2_bfrag<H, false>(t, gid, tig, [&](int h, int g) { return __ldg(wl + g * H + h) * (scale * kLog2e); });
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int c = 0; c < KS; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) split_tf32(__uint_as_float(t[a][c][e]), b1h[a][c][e], b1l[a][c][e]);
        th2_bfrag<H, true>(b2, gid, tig, [&](int g, int g2) { return __ldg(ww + g2 * H + g); });
    }
    float bl2[KS][2], bw2[KS][2];   // biases of the heads this lane holds in C layout
#pragma unroll
    for (int nt = 0; nt < KS; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int g = 8 * nt + 2 * tig + e;
            bl2[nt][e] = g < H ? __ldg(bl + g) * kLog2e : 0.f;
            bw2[nt][e] = g < H ? __ldg(bw + g) : 0.f;
        }

    const long long rows = (long long)B * N;
    for (long long row = (long long)blockIdx.x * TH2_WARPS + warp; row < rows; row += (long long)gridDim.x * TH2_WARPS) {
        const int b = static_cast<int>(row / N), i = static_cast<int>(row - (long long)b * N);
        const long long base = ((long long)b * H * N + i) * Np;
        const ST* Srow = S + base;
        // all loads of the row go out first (NT tiles x 2 x KS loads per lane in flight): the row is one long
        // dependent chain otherwise and the kernel would run at one HBM round trip per tile
        RawS raw[NT][KS][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t < nfull) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        raw[t][ks][e] = off.a_ok[ks][e] ? th2_load_raw(Srow + off.a[ks][e] + t * 16) : RawS();
            } else {
                const bool inb = t * 16 + 2 * gid < Np;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        raw[t][ks][e] = (inb && off.a_ok[ks][e]) ? th2_load_raw(Srow + off.a[ks][e] + t * 16) : RawS();
            }
        }
        float sp[NT][KS][4];   // mixed logits (log2 domain), C layout
        float mx[KS][2];
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) mx[nt][0] = mx[nt][1] = -INFINITY;
        auto logits_tile = [&](auto masked_c, int t) {
            constexpr bool MASKED = decltype(masked_c)::value;
            uint32_t ah[KS][4], al[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float2 v0 = th2_raw_f2(raw[t][ks][0]), v1 = th2_raw_f2(raw[t][ks][1]);
                if constexpr (sizeof(ST) == 4) {
                    split_tf32(v0.x, ah[ks][0], al[ks][0]);
                    split_tf32(v0.y, ah[ks][1], al[ks][1]);
                    split_tf32(v1.x, ah[ks][2], al[ks][2]);
                    split_tf32(v1.y, ah[ks][3], al[ks][3]);
                } else {    // bf16 values are tf32 values: no rounding, no low part
                    ah[ks][0] = __float_as_uint(v0.x); ah[ks][1] = __float_as_uint(v0.y);
                    ah[ks][2] = __float_as_uint(v1.x); ah[ks][3] = __float_as_uint(v1.y);
                    al[ks][0] = al[ks][1] = al[ks][2] = al[ks][3] = 0u;
                }
            }
#pragma unroll
            for (int nt = 0; nt < KS; ++nt) {
                float d[4] = {bl2[nt][0], bl2[nt][1], bl2[nt][0], bl2[nt][1]};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    if constexpr (sizeof(ST) == 4) mma_tf32(d, al[ks], b1h[ks][nt]);
                    mma_tf32(d, ah[ks], b1l[ks][nt]);
                    mma_tf32(d, ah[ks], b1h[ks][nt]);
                }
                if constexpr (MASKED) {
                    const int col = t * 16 + 2 * gid;
                    if (col >= N) d[0] = d[1] = -INFINITY;          // columns 2gid / 2gid+1 of the tile
                    if (col + 1 >= N) d[2] = d[3] = -INFINITY;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) sp[t][nt][e] = d[e];
                mx[nt][0] = fmaxf(mx[nt][0], fmaxf(d[0], d[2]));
                mx[nt][1] = fmaxf(mx[nt][1], fmaxf(d[1], d[3]));
            }
        };
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t < nfull) logits_tile(Th2False(), t);
            else logits_tile(Th2True(), t);
        }
        float sum[KS][2], inv[KS][2];
#pragma unroll
        for (int nt = 0; nt < KS; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float m = mx[nt][e];
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                mx[nt][e] = m;
                sum[nt][e] = 0.f;
            }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            {
#pragma unroll
                for (int nt = 0; nt < KS; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p = ex2_approx(sp[t][nt][e] - mx[nt][e & 1]);   // exp2(-inf) = 0 for masked columns
                        sp[t][nt][e] = p;
                        sum[nt][e & 1] += p;
                    }
            }
        }
#pragma unroll
        for (int nt = 0; nt < KS; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float s = sum[nt][e];
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                sum[nt][e] = s;
                inv[nt][e] = 1.0f / s;
                const int g = 8 * nt + 2 * tig + e;
                if (gid == 0 && g < H) {    // saved for backward in the natural-log domain of the forward definition
                    rowmax[((long long)b * H + g) * N + i] = mx[nt][e] * 0.6931471805599453f;
                    rowsum[((long long)b * H + g) * N + i] = s;
                }
            }
        __nv_bfloat16* Prow = Pm + base;
        auto probs_tile = [&](auto masked_c, int t) {
            constexpr bool MASKED = decltype(masked_c)::value;
            uint32_t a[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                float c[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) c[e] = sp[t][ks][e] * inv[ks][e & 1];
                th2_c_to_a(c, a[ks]);
            }
#pragma unroll
            for (int nt = 0; nt < KS; ++nt) {
                float d[4] = {bw2[nt][0], bw2[nt][1], bw2[nt][0], bw2[nt][1]};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) mma_tf32(d, a[ks], b2[ks][nt]);
                if constexpr (MASKED) {
                    const int col = t * 16 + 2 * gid;
                    if (col < Np) {
                        const float v00 = col < N ? d[0] : 0.f, v01 = col + 1 < N ? d[2] : 0.f;   // head g0, cols col, col+1
                        const float v10 = col < N ? d[1] : 0.f, v11 = col + 1 < N ? d[3] : 0.f;   // head g0+1
                        if (off.c_ok[nt][0]) *reinterpret_cast<uint32_t*>(Prow + off.c[nt][0] + t * 16) = pack_bf16(v00, v01);
                        if (off.c_ok[nt][1]) *reinterpret_cast<uint32_t*>(Prow + off.c[nt][1] + t * 16) = pack_bf16(v10, v11);
                    }
                } else {
                    if (off.c_ok[nt][0]) *reinterpret_cast<uint32_t*>(Prow + off.c[nt][0] + t * 16) = pack_bf16(d[0], d[2]);
                    if (off.c_ok[nt][1]) *reinterpret_cast<uint32_t*>(Prow + off.c[nt][1] + t * 16) = pack_bf16(d[1], d[3]);
                }
            }
        };
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t < nfull) probs_tile(Th2False(), t);
            else probs_tile(Th2True(), t);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
template <int H> struct Th2Scratch {
    static constexpr int HS = 20;                                 // floats between heads: conflict-free transposes
    static constexpr int ARR = Th2<H>::HP * HS;                   // one [head][16 columns] array
    static constexpr int FLOATS = 4 * ARR;                        // dP' | dS' | P | S
};

// bf16 logit planes run the LEAN form: the row stays in the (packed) registers it was loaded into and phase 2 recomputes
// P and dP of a tile with three more mma.sync instead of keeping 8 fp32 per tile per lane alive across the row -- half
// the registers, two thread blocks per SM (the kernel is bound by the latency of its dependent mma / exp chains, not by
// the tensor pipe: 8 clk per mma.sync per SM sub-partition, scripts/micro/pipe_rates.cu).
// Variants (WARPS per block, MINB blocks per SM, LEAN): (8, 1, false) the original; (8, 2, true) 128 registers; (12, 1,
// true) 168 registers. VITK_TH_BWD picks one at run time (cait_attn.cu) for A/B runs.
template <int H, int NT, typename ST, int WARPS, int MINB, bool LEAN_>
__global__ void __launch_bounds__(WARPS * 32, MINB)
th_mix2_bwd_kernel(const ST* __restrict__ S, const __nv_bfloat16* __restrict__ dPm, const float* __restrict__ rowmax,
                   const float* __restrict__ rowsum, const float* __restrict__ wl, const float* __restrict__ bl,
                   const float* __restrict__ ww, float scale, __nv_bfloat16* __restrict__ dS, float* __restrict__ dwl,
                   float* __restrict__ dbl, float* __restrict__ dww, float* __restrict__ dbw, int B, int N, int Np) {
    constexpr int KS = Th2<H>::KS;
    constexpr int HS = Th2Scratch<H>::HS, ARR = Th2Scratch<H>::ARR;
    __shared__ __align__(16) float scratch_all[WARPS][Th2Scratch<H>::FLOATS];
    __shared__ float red[2 * H * H + 2 * H];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long plane = (long long)N * Np;
    constexpr float kLog2e = 1.4426950408889634f;
    float* sc_dpm = scratch_all[warp];
    float* sc_dsp = sc_dpm + ARR;
    float* sc_p = sc_dsp + ARR;
    float* sc_s = sc_p + ARR;
    for (int k = threadIdx.x; k < 2 * H * H + 2 * H; k += blockDim.x) red[k] = 0.f;
    __syncthreads();

    uint32_t b1h[KS][KS][2], b1l[KS][KS][2], bdp[KS][KS][2], bds[KS][KS][2];
    {
        uint32_t t[KS][KS][2];
        th2_bfrag<H, false>(t, gid, tig, [&](int h, int g) { return __ldg(wl + g * H + h) * (scale * kLog2e); });
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int c = 0; c < KS; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) split_tf32(__uint_as_float(t[a][c][e]), b1h[a][c][e], b1l[a][c][e]);
        // dP[col][g] = sum_g' dP'[col][g'] Ww[g'][g]     (A from memory: natural K slots)
        th2_bfrag<H, false>(bdp, gid, tig, [&](int g2, int g) { return __ldg(ww + g2 * H + g); });
        // dS[col][h] = sum_g dS'[col][g] scale*Wl[g][h]   (A = C fragment: permuted K slots)
        th2_bfrag<H, true>(bds, gid, tig, [&](int g, int h) { return __ldg(wl + g * H + h) * scale; });
    }
    float bl2[KS][2];
#pragma unroll
    for (int nt = 0; nt < KS; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int g = 8 * nt + 2 * tig + e;
            bl2[nt][e] = g < H ? __ldg(bl + g) * kLog2e : 0.f;
        }
    // weight-gradient accumulators: G1[g'][g] = sum dP'[g'] P[g] ; G2[g][h] = sum dS'[g] S[h]  (rows = MMA M, 16 heads)
    float G1[KS][4], G2[KS][4];
#pragma unroll
    for (int nt = 0; nt < KS; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) G1[nt][e] = G2[nt][e] = 0.f;
    float acc_dbw[KS][2], acc_dbl[KS][2];   // dbw: heads tig / tig+4 (A layout) ; dbl: heads 2tig / 2tig+1 (C layout)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) acc_dbw[ks][0] = acc_dbw[ks][1] = acc_dbl[ks][0] = acc_dbl[ks][1] = 0.f;

    const long long rows = (long long)B * N;
    for (long long row = (long long)blockIdx.x * WARPS + warp; row < rows; row += (long long)gridDim.x * WARPS) {
        const int b = static_cast<int>(row / N), i = static_cast<int>(row - (long long)b * N);
        const long long base = ((long long)b * H * N + i) * Np;
        float m2[KS][2], inv[KS][2];
#pragma unroll
        for (int nt = 0; nt < KS; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int g = 8 * nt + 2 * tig + e;
                m2[nt][e] = g < H ? rowmax[((long long)b * H + g) * N + i] * kLog2e : 0.f;
                inv[nt][e] = g < H ? 1.0f / rowsum[((long long)b * H + g) * N + i] : 0.f;
            }
        // all loads of the row first (see the forward kernel)
        using RawS = typename Th2Raw<ST>::type;
        constexpr bool LEAN = LEAN_;
        RawS rs[NT][KS][2];
        uint32_t rd[NT][KS][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int col = t * 16 + 2 * gid;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int h0 = 8 * ks + tig, h1 = h0 + 4;
                const bool ok0 = col < Np && h0 < H, ok1 = col < Np && h1 < H;
                rs[t][ks][0] = ok0 ? th2_load_raw(S + base + h0 * plane + col) : RawS();
                rs[t][ks][1] = ok1 ? th2_load_raw(S + base + h1 * plane + col) : RawS();
                rd[t][ks][0] = ok0 ? __ldg(reinterpret_cast<const uint32_t*>(dPm + base + h0 * plane + col)) : 0u;
                rd[t][ks][1] = ok1 ? __ldg(reinterpret_cast<const uint32_t*>(dPm + base + h1 * plane + col)) : 0u;
            }
        }
        float P[LEAN ? 1 : NT][KS][4], dP[LEAN ? 1 : NT][KS][4];
        float rp[KS][2];
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) rp[nt][0] = rp[nt][1] = 0.f;
        // P and dP of tile t (C layout) from the row registers
        auto tile_pd = [&](int t, float (&pt)[KS][4], float (&qt)[KS][4]) {
            const int col = t * 16 + 2 * gid;
            uint32_t ah[KS][4], al[KS][4], ad[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float2 v0 = th2_raw_f2(rs[t][ks][0]), v1 = th2_raw_f2(rs[t][ks][1]);
                if constexpr (sizeof(ST) == 4) {
                    split_tf32(v0.x, ah[ks][0], al[ks][0]);
                    split_tf32(v0.y, ah[ks][1], al[ks][1]);
                    split_tf32(v1.x, ah[ks][2], al[ks][2]);
                    split_tf32(v1.y, ah[ks][3], al[ks][3]);
                } else {    // bf16 values are tf32 values: no rounding, no low part
                    ah[ks][0] = __float_as_uint(v0.x); ah[ks][1] = __float_as_uint(v0.y);
                    ah[ks][2] = __float_as_uint(v1.x); ah[ks][3] = __float_as_uint(v1.y);
                    al[ks][0] = al[ks][1] = al[ks][2] = al[ks][3] = 0u;
                }
                const uint32_t d0 = rd[t][ks][0], d1 = rd[t][ks][1];
                ad[ks][0] = col < N ? (d0 << 16) : 0u;
                ad[ks][1] = col + 1 < N ? (d0 & 0xffff0000u) : 0u;
                ad[ks][2] = col < N ? (d1 << 16) : 0u;
                ad[ks][3] = col + 1 < N ? (d1 & 0xffff0000u) : 0u;
            }
#pragma unroll
            for (int nt = 0; nt < KS; ++nt) {
                float d[4] = {bl2[nt][0], bl2[nt][1], bl2[nt][0], bl2[nt][1]};
                float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    if constexpr (sizeof(ST) == 4) mma_tf32(d, al[ks], b1h[ks][nt]);
                    mma_tf32(d, ah[ks], b1l[ks][nt]);
                    mma_tf32(d, ah[ks], b1h[ks][nt]);
                    mma_tf32(q, ad[ks], bdp[ks][nt]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool valid = (col + (e >> 1)) < N;
                    pt[nt][e] = valid ? ex2_approx(d[e] - m2[nt][e & 1]) * inv[nt][e & 1] : 0.f;
                    qt[nt][e] = q[e];
                }
            }
        };
        // ---- phase 1: P, dP and rowsum(dP o P)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if constexpr (LEAN) {
                float pt[KS][4], qt[KS][4];
                tile_pd(t, pt, qt);
#pragma unroll
                for (int nt = 0; nt < KS; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) rp[nt][e & 1] = fmaf(pt[nt][e], qt[nt][e], rp[nt][e & 1]);
            } else {
                const int col = t * 16 + 2 * gid;
                uint32_t ah[KS][4], al[KS][4], ad[KS][4];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const float2 v0 = th2_raw_f2(rs[t][ks][0]), v1 = th2_raw_f2(rs[t][ks][1]);
                    split_tf32(v0.x, ah[ks][0], al[ks][0]);
                    split_tf32(v0.y, ah[ks][1], al[ks][1]);
                    split_tf32(v1.x, ah[ks][2], al[ks][2]);
                    split_tf32(v1.y, ah[ks][3], al[ks][3]);
                    const uint32_t d0 = rd[t][ks][0], d1 = rd[t][ks][1];
                    // bf16 -> fp32 bit patterns (exactly representable in tf32); the pad columns [N, Np) of the buffers
                    // are never written by the producing GEMM and may hold anything: zero them
                    ad[ks][0] = col < N ? (d0 << 16) : 0u;
                    ad[ks][1] = col + 1 < N ? (d0 & 0xffff0000u) : 0u;
                    ad[ks][2] = col < N ? (d1 << 16) : 0u;
                    ad[ks][3] = col + 1 < N ? (d1 & 0xffff0000u) : 0u;
                }
#pragma unroll
                for (int nt = 0; nt < KS; ++nt) {
                    float d[4] = {bl2[nt][0], bl2[nt][1], bl2[nt][0], bl2[nt][1]};
                    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        if constexpr (sizeof(ST) == 4) mma_tf32(d, al[ks], b1h[ks][nt]);
                        mma_tf32(d, ah[ks], b1l[ks][nt]);
                        mma_tf32(d, ah[ks], b1h[ks][nt]);
                        mma_tf32(q, ad[ks], bdp[ks][nt]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool valid = (col + (e >> 1)) < N;
                        const float p = valid ? ex2_approx(d[e] - m2[nt][e & 1]) * inv[nt][e & 1] : 0.f;
                        P[t][nt][e] = p;
                        dP[t][nt][e] = q[e];
                        rp[nt][e & 1] = fmaf(p, q[e], rp[nt][e & 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < KS; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float r = rp[nt][e];
                r += __shfl_xor_sync(0xffffffffu, r, 4);
                r += __shfl_xor_sync(0xffffffffu, r, 8);
                r += __shfl_xor_sync(0xffffffffu, r, 16);
                rp[nt][e] = r;
            }
        if constexpr (LEAN) {
            // phase 2 recomputes from the row registers: keep the compiler from carrying phase 1's operand fragments
            // across the row instead (that would be the register footprint this form exists to avoid)
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        asm volatile("" : "+r"(rd[t][ks][e]));
                        if constexpr (sizeof(RawS) == 4) asm volatile("" : "+r"(*reinterpret_cast<uint32_t*>(&rs[t][ks][e])));
                    }
        }
        // ---- phase 2: dS', dS, weight gradients
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            {
                const int col = t * 16 + 2 * gid;
                const bool inb = col < Np;
                float dsp[KS][4];
                uint32_t a[KS][4];
                float pt[KS][4], qt[KS][4];
                if constexpr (LEAN) {
                    tile_pd(t, pt, qt);
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            pt[ks][e] = P[t][ks][e];
                            qt[ks][e] = dP[t][ks][e];
                        }
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) dsp[ks][e] = pt[ks][e] * (qt[ks][e] - rp[ks][e & 1]);
                    th2_c_to_a(dsp[ks], a[ks]);
                    acc_dbl[ks][0] += dsp[ks][0] + dsp[ks][2];
                    acc_dbl[ks][1] += dsp[ks][1] + dsp[ks][3];
                }
#pragma unroll
                for (int nt = 0; nt < KS; ++nt) {
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) mma_tf32(d, a[ks], bds[ks][nt]);
                    if (inb) {
                        const int h0 = 8 * nt + 2 * tig;
                        const float v00 = col < N ? d[0] : 0.f, v01 = col + 1 < N ? d[2] : 0.f;
                        const float v10 = col < N ? d[1] : 0.f, v11 = col + 1 < N ? d[3] : 0.f;
                        if (h0 < H) *reinterpret_cast<uint32_t*>(dS + base + h0 * plane + col) = pack_bf16(v00, v01);
                        if (h0 + 1 < H) *reinterpret_cast<uint32_t*>(dS + base + (h0 + 1) * plane + col) = pack_bf16(v10, v11);
                    }
                }
                // transposes through the per-warp scratch: [head][16 columns]
                __syncwarp();
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int h0 = 8 * ks + tig, h1 = h0 + 4;
                    // S and dP' in A layout: still in the registers the row was loaded into
                    const float2 s0 = th2_raw_f2(rs[t][ks][0]), s1 = th2_raw_f2(rs[t][ks][1]);
                    const uint32_t d0 = rd[t][ks][0], d1 = rd[t][ks][1];
                    const float2 e0 = make_float2(col < N ? bf16_lo(d0) : 0.f, col + 1 < N ? bf16_hi(d0) : 0.f);
                    const float2 e1 = make_float2(col < N ? bf16_lo(d1) : 0.f, col + 1 < N ? bf16_hi(d1) : 0.f);
                    const float2 t0 = make_float2(col < N ? s0.x : 0.f, col + 1 < N ? s0.y : 0.f);
                    const float2 t1 = make_float2(col < N ? s1.x : 0.f, col + 1 < N ? s1.y : 0.f);
                    acc_dbw[ks][0] += e0.x + e0.y;
                    acc_dbw[ks][1] += e1.x + e1.y;
                    *reinterpret_cast<float2*>(sc_s + h0 * HS + 2 * gid) = t0;
                    *reinterpret_cast<float2*>(sc_s + h1 * HS + 2 * gid) = t1;
                    *reinterpret_cast<float2*>(sc_dpm + h0 * HS + 2 * gid) = e0;
                    *reinterpret_cast<float2*>(sc_dpm + h1 * HS + 2 * gid) = e1;
                    const int g0 = 8 * ks + 2 * tig;   // C layout: heads g0, g0+1
                    *reinterpret_cast<float2*>(sc_p + g0 * HS + 2 * gid) = make_float2(pt[ks][0], pt[ks][2]);
                    *reinterpret_cast<float2*>(sc_p + (g0 + 1) * HS + 2 * gid) = make_float2(pt[ks][1], pt[ks][3]);
                    *reinterpret_cast<float2*>(sc_dsp + g0 * HS + 2 * gid) = make_float2(dsp[ks][0], dsp[ks][2]);
                    *reinterpret_cast<float2*>(sc_dsp + (g0 + 1) * HS + 2 * gid) = make_float2(dsp[ks][1], dsp[ks][3]);
                }
                __syncwarp();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {     // two K steps of 8 columns
                    const int c0 = 8 * kk + tig, c1 = c0 + 4;
                    uint32_t a1[4], a2[4];           // A[m = head][k = column]
                    a1[0] = __float_as_uint(sc_dpm[gid * HS + c0]);
                    a1[2] = __float_as_uint(sc_dpm[gid * HS + c1]);
                    a2[0] = __float_as_uint(sc_dsp[gid * HS + c0]);
                    a2[2] = __float_as_uint(sc_dsp[gid * HS + c1]);
                    if (KS > 1) {
                        a1[1] = __float_as_uint(sc_dpm[(gid + 8) * HS + c0]);
                        a1[3] = __float_as_uint(sc_dpm[(gid + 8) * HS + c1]);
                        a2[1] = __float_as_uint(sc_dsp[(gid + 8) * HS + c0]);
                        a2[3] = __float_as_uint(sc_dsp[(gid + 8) * HS + c1]);
                    } else {
                        a1[1] = a1[3] = a2[1] = a2[3] = 0u;
                    }
#pragma unroll
                    for (int nt = 0; nt < KS; ++nt) {
                        uint32_t bp[2], bs[2];       // B[k = column][n = head]
                        bp[0] = __float_as_uint(sc_p[(8 * nt + gid) * HS + c0]);
                        bp[1] = __float_as_uint(sc_p[(8 * nt + gid) * HS + c1]);
                        bs[0] = __float_as_uint(sc_s[(8 * nt + gid) * HS + c0]);
                        bs[1] = __float_as_uint(sc_s[(8 * nt + gid) * HS + c1]);
                        mma_tf32(G1[nt], a1, bp);
                        mma_tf32(G2[nt], a2, bs);
                    }
                }
            }
        }
    }
    // ---- reduce the weight gradients over the block's warps, then one atomic per entry
    // G C-layout: c0 = (row gid, col 2tig), c1 = (gid, 2tig+1), c2 = (gid+8, 2tig), c3 = (gid+8, 2tig+1), col += 8nt
#pragma unroll
    for (int nt = 0; nt < KS; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = gid + 8 * (e >> 1), c = 8 * nt + 2 * tig + (e & 1);
            if (r < H && c < H) {
                atomicAdd(&red[r * H + c], G1[nt][e]);                    // dWw[g'][g]
                atomicAdd(&red[H * H + r * H + c], G2[nt][e] * scale);    // dWl[g][h]
            }
        }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float v = acc_dbw[ks][e];     // head 8ks + tig + 4e, summed over the lanes that share tig
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            const int g = 8 * ks + tig + 4 * e;
            if (gid == 0 && g < H) atomicAdd(&red[2 * H * H + g], v);
            float u = acc_dbl[ks][e];     // head 8ks + 2tig + e
            u += __shfl_xor_sync(0xffffffffu, u, 4);
            u += __shfl_xor_sync(0xffffffffu, u, 8);
            u += __shfl_xor_sync(0xffffffffu, u, 16);
            const int g2 = 8 * ks + 2 * tig + e;
            if (gid == 0 && g2 < H) atomicAdd(&red[2 * H * H + H + g2], u);
        }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * H * H + 2 * H; k += blockDim.x) {
        const float v = red[k];
        if (k < H * H) atomicAdd(dww + k, v);
        else if (k < 2 * H * H) atomicAdd(dwl + (k - H * H), v);
        else if (k < 2 * H * H + H) atomicAdd(dbw + (k - 2 * H * H), v);
        else atomicAdd(dbl + (k - 2 * H * H - H), v);
    }
}

}  // namespace vitk
