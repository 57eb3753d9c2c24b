// CaiT-specific attention glue (memory-bound CUDA-core kernels; the tensor-core products around them are batched
// tcgen05 GEMMs, see vitk_gemm_bf16_batched):
//
//  * talking-heads mixing (models/cait.py:116-125):  S' = Wl S + bl over the head axis, P = softmax_j(S'),
//    P' = Ww P + bw. th_mix_fwd reads the fp32 logits S[B,H,N,Np] once and writes the bf16 mixed probabilities
//    P'[B,H,N,Np] once (the eager reference makes ~8 fp32 passes over [B,H,N,N] incl. two permute copies).
//    th_mix_bwd turns dP' into dS and accumulates dWl, dbl, dWw, dbw.
//  * class attention (models/cait.py:38-55): one query row (the class token) per (image, head): warp-level
//    GEMV - softmax - GEMV, forward and backward.
#include "common.cuh"
#include "tmap.cuh"
#include "th_mix2.cuh"
#include "../../include/vitk.h"
#include <stdlib.h>

namespace vitk {

constexpr int TH_THREADS = 128;

template <int H> struct ThWeights {
    float wl[H * H];  // scale * Wl[g][h]
    float bl[H];
    float ww[H * H];  // Ww[g][h]
    float bw[H];
};

template <int H>
__device__ __forceinline__ void th_load_weights(ThWeights<H>* sw, const float* wl, const float* bl, const float* ww,
                                                const float* bw, float scale) {
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
        sw->wl[i] = wl[i] * scale;
        sw->ww[i] = ww[i];
    }
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        sw->bl[i] = bl[i];
        sw->bw[i] = bw[i];
    }
}

// block-wide reduction of H per-thread values (max or sum); result broadcast to every thread
template <int H, bool IS_MAX>
__device__ __forceinline__ void th_block_reduce(float* v, float (*red)[H]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < H; ++g) {
        const float r = IS_MAX ? warp_max(v[g]) : warp_sum(v[g]);
        if (lane == 0) red[warp][g] = r;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < H; ++g) {
        float r = red[0][g];
#pragma unroll
        for (int w = 1; w < TH_THREADS / 32; ++w) r = IS_MAX ? fmaxf(r, red[w][g]) : r + red[w][g];
        v[g] = r;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// forward: one CTA per (b, i) row; thread t owns columns j = t + 128*jj
// ------------------------------------------------------------------------------------------------------------------
template <int H, int JT>
__global__ void __launch_bounds__(TH_THREADS)
th_mix_fwd_kernel(const float* __restrict__ S, const float* __restrict__ wl, const float* __restrict__ bl,
                  const float* __restrict__ ww, const float* __restrict__ bw, float scale, __nv_bfloat16* __restrict__ Pm,
                  float* __restrict__ rowmax, float* __restrict__ rowsum, int B, int N, int Np) {
    __shared__ ThWeights<H> sw;
    __shared__ float red[TH_THREADS / 32][H];
    th_load_weights<H>(&sw, wl, bl, ww, bw, scale);
    __syncthreads();
    const long long row = blockIdx.x;  // b * N + i
    const int b = static_cast<int>(row / N), i = static_cast<int>(row - (long long)b * N);
    const long long plane = (long long)N * Np;
    const float* Srow = S + ((long long)b * H * N + i) * Np;
    float sp[JT][H];
    float mx[H];
#pragma unroll
    for (int g = 0; g < H; ++g) mx[g] = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) {
        const int j = threadIdx.x + jj * TH_THREADS;
        const bool valid = j < N;
        float s[H];
#pragma unroll
        for (int h = 0; h < H; ++h) s[h] = valid ? Srow[h * plane + j] : 0.f;
#pragma unroll
        for (int g = 0; g < H; ++g) {
            float a = sw.bl[g];
#pragma unroll
            for (int h = 0; h < H; ++h) a = fmaf(sw.wl[g * H + h], s[h], a);
            sp[jj][g] = valid ? a : -INFINITY;
            mx[g] = fmaxf(mx[g], sp[jj][g]);
        }
    }
    th_block_reduce<H, true>(mx, red);
    float sum[H];
#pragma unroll
    for (int g = 0; g < H; ++g) sum[g] = 0.f;
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
#pragma unroll
        for (int g = 0; g < H; ++g) {
            sp[jj][g] = __expf(sp[jj][g] - mx[g]);  // exp(-inf) = 0 for padded columns
            sum[g] += sp[jj][g];
        }
    th_block_reduce<H, false>(sum, red);
    if (threadIdx.x < H) {
        rowmax[((long long)b * H + threadIdx.x) * N + i] = mx[threadIdx.x];
        rowsum[((long long)b * H + threadIdx.x) * N + i] = sum[threadIdx.x];
    }
    float inv[H];
#pragma unroll
    for (int g = 0; g < H; ++g) inv[g] = 1.0f / sum[g];
    __nv_bfloat16* Prow = Pm + ((long long)b * H * N + i) * Np;
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) {
        const int j = threadIdx.x + jj * TH_THREADS;
        if (j < Np) {
            const bool valid = j < N;
#pragma unroll
            for (int g2 = 0; g2 < H; ++g2) {
                float a = sw.bw[g2];
#pragma unroll
                for (int g = 0; g < H; ++g) a = fmaf(sw.ww[g2 * H + g], sp[jj][g] * inv[g], a);
                Prow[g2 * plane + j] = __float2bfloat16_rn(valid ? a : 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: persistent CTAs over (b, i) rows.
//   dp_h = sum_g Ww[g,h] dP'_g ; dS'_h = P_h (dp_h - sum_j dp_h P_h) ; dS_h = sum_g scale*Wl[g,h] dS'_g
//   dWw[g,h] += sum dP'_g P_h ; dbw[g] += sum dP'_g ; dWl[g,h] += scale * sum dS'_g S_h ; dbl[g] += sum dS'_g
// ------------------------------------------------------------------------------------------------------------------
template <int H, int JT>
__global__ void __launch_bounds__(TH_THREADS)
th_mix_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dPm, const float* __restrict__ rowmax,
                  const float* __restrict__ rowsum, const float* __restrict__ wl, const float* __restrict__ bl,
                  const float* __restrict__ ww, const float* __restrict__ bw, float scale,
                  __nv_bfloat16* __restrict__ dS, float* __restrict__ dwl, float* __restrict__ dbl,
                  float* __restrict__ dww, float* __restrict__ dbw, int B, int N, int Np) {
    extern __shared__ __align__(16) float th_smem[];
    __shared__ ThWeights<H> sw;
    __shared__ float red[TH_THREADS / 32][H];
    const int Ns = Np + 1;  // odd-ish pitch: rows of different heads fall into different banks
    float* sm_s = th_smem;                 // [H][Ns] logits
    float* sm_dpm = sm_s + H * Ns;         // [H][Ns] dP'
    float* sm_p = sm_dpm + H * Ns;         // [H][Ns] softmax probabilities
    float* sm_dsp = sm_p + H * Ns;         // [H][Ns] dS'
    th_load_weights<H>(&sw, wl, bl, ww, bw, scale);
    __syncthreads();
    constexpr int NITEMS = 2 * H * H + 2 * H;
    constexpr int IPT = (NITEMS + TH_THREADS - 1) / TH_THREADS;
    float acc[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) acc[k] = 0.f;
    const long long plane = (long long)N * Np;
    const long long rows = (long long)B * N;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int b = static_cast<int>(row / N), i = static_cast<int>(row - (long long)b * N);
        const long long base = ((long long)b * H * N + i) * Np;
        float mrow[H], irow[H];
#pragma unroll
        for (int g = 0; g < H; ++g) {
            mrow[g] = rowmax[((long long)b * H + g) * N + i];
            irow[g] = 1.0f / rowsum[((long long)b * H + g) * N + i];
        }
        float p[JT][H], dp[JT][H], rpart[H];
#pragma unroll
        for (int h = 0; h < H; ++h) rpart[h] = 0.f;
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            const int j = threadIdx.x + jj * TH_THREADS;
            const bool valid = j < N;
            float s[H], d[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                s[h] = valid ? S[base + h * plane + j] : 0.f;
                d[h] = valid ? dPm[base + h * plane + j] : 0.f;
                if (j < Np) {
                    sm_s[h * Ns + j] = s[h];
                    sm_dpm[h * Ns + j] = d[h];
                }
            }
#pragma unroll
            for (int g = 0; g < H; ++g) {
                float a = sw.bl[g];
#pragma unroll
                for (int h = 0; h < H; ++h) a = fmaf(sw.wl[g * H + h], s[h], a);
                p[jj][g] = valid ? __expf(a - mrow[g]) * irow[g] : 0.f;
                if (j < Np) sm_p[g * Ns + j] = p[jj][g];
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float a = 0.f;
#pragma unroll
                for (int g = 0; g < H; ++g) a = fmaf(sw.ww[g * H + h], d[g], a);
                dp[jj][h] = a;
                rpart[h] = fmaf(a, p[jj][h], rpart[h]);
            }
        }
        th_block_reduce<H, false>(rpart, red);
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            const int j = threadIdx.x + jj * TH_THREADS;
            if (j < Np) {
                float dsp[H];
#pragma unroll
                for (int g = 0; g < H; ++g) {
                    dsp[g] = p[jj][g] * (dp[jj][g] - rpart[g]);
                    sm_dsp[g * Ns + j] = dsp[g];
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float a = 0.f;
#pragma unroll
                    for (int g = 0; g < H; ++g) a = fmaf(sw.wl[g * H + h], dsp[g], a);
                    dS[base + h * plane + j] = __float2bfloat16_rn(j < N ? a : 0.f);
                }
            }
        }
        __syncthreads();
        // weight-gradient dot products over the row held in shared memory
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int item = threadIdx.x + k * TH_THREADS;
            if (item < NITEMS) {
                const float* u;
                const float* v = nullptr;
                if (item < H * H) {                       // dWw[g][h] : dP'_g . P_h
                    u = sm_dpm + (item / H) * Ns;
                    v = sm_p + (item % H) * Ns;
                } else if (item < 2 * H * H) {            // dWl[g][h] : dS'_g . S_h (x scale at the end)
                    const int it = item - H * H;
                    u = sm_dsp + (it / H) * Ns;
                    v = sm_s + (it % H) * Ns;
                } else if (item < 2 * H * H + H) {        // dbw[g]
                    u = sm_dpm + (item - 2 * H * H) * Ns;
                } else {                                  // dbl[g]
                    u = sm_dsp + (item - 2 * H * H - H) * Ns;
                }
                float a = 0.f;
                if (v != nullptr) {
                    for (int j = 0; j < N; ++j) a = fmaf(u[j], v[j], a);
                } else {
                    for (int j = 0; j < N; ++j) a += u[j];
                }
                acc[k] += a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int item = threadIdx.x + k * TH_THREADS;
        if (item < H * H) atomicAdd(dww + item, acc[k]);
        else if (item < 2 * H * H) atomicAdd(dwl + (item - H * H), acc[k] * scale);
        else if (item < 2 * H * H + H) atomicAdd(dbw + (item - 2 * H * H), acc[k]);
        else if (item < NITEMS) atomicAdd(dbl + (item - 2 * H * H - H), acc[k]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// class attention: one warp per (b, h). Keys/values: row 0 = class token (kc/vc [B, C]), rows 1..n = patch tokens
// (kx/vx [B*n, ldkv]); q [B, C] unscaled. p (softmax probabilities, fp32 [B,H,n+1]) is saved for backward.
// ------------------------------------------------------------------------------------------------------------------
constexpr int CA_WARPS = 4;

template <int HD>
__global__ void __launch_bounds__(CA_WARPS * 32)
class_attn_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                      const __nv_bfloat16* __restrict__ kx, const __nv_bfloat16* __restrict__ vc,
                      const __nv_bfloat16* __restrict__ vx, long long ldkv, long long ldc, float scale,
                      __nv_bfloat16* __restrict__ out, float* __restrict__ p_out, int B, int H, int n) {
    extern __shared__ __align__(16) float ca_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = n + 1, C = H * HD;
    float* sq = ca_smem + warp * (HD + N);  // [HD] scaled query, then [N] scores / probabilities
    float* sp = sq + HD;
    const int bh = blockIdx.x * CA_WARPS + warp;
    if (bh >= B * H) return;
    const int b = bh / H, h = bh - b * H;
    for (int e = lane; e < HD; e += 32) sq[e] = __bfloat162float(q[(long long)b * C + h * HD + e]) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
        const __nv_bfloat16* kr = (j == 0) ? kc + (long long)b * ldc + h * HD : kx + ((long long)b * n + j - 1) * ldkv + h * HD;
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
            const uint4 t = *reinterpret_cast<const uint4*>(kr + u * 8);
            a += sq[u * 8 + 0] * bf16_lo(t.x) + sq[u * 8 + 1] * bf16_hi(t.x) + sq[u * 8 + 2] * bf16_lo(t.y) +
                 sq[u * 8 + 3] * bf16_hi(t.y) + sq[u * 8 + 4] * bf16_lo(t.z) + sq[u * 8 + 5] * bf16_hi(t.z) +
                 sq[u * 8 + 6] * bf16_lo(t.w) + sq[u * 8 + 7] * bf16_hi(t.w);
        }
        sp[j] = a;
        mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float e = __expf(sp[j] - mx);
        sp[j] = e;
        sum += e;
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < N; j += 32) {
        sp[j] *= inv;
        p_out[(long long)bh * N + j] = sp[j];
    }
    __syncwarp();
    // out[e] = sum_j p_j v[j][e]; each lane owns two adjacent channels
    const int e0 = lane * 2;
    if (e0 < HD) {
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < N; ++j) {
            const __nv_bfloat16* vr = (j == 0) ? vc + (long long)b * ldc + h * HD : vx + ((long long)b * n + j - 1) * ldkv + h * HD;
            const uint32_t t = *reinterpret_cast<const uint32_t*>(vr + e0);
            a0 = fmaf(sp[j], bf16_lo(t), a0);
            a1 = fmaf(sp[j], bf16_hi(t), a1);
        }
        *reinterpret_cast<uint32_t*>(out + (long long)b * C + h * HD + e0) = pack_bf16(a0, a1);
    }
}

template <int HD>
__global__ void __launch_bounds__(CA_WARPS * 32)
class_attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                      const __nv_bfloat16* __restrict__ kx, const __nv_bfloat16* __restrict__ vc,
                      const __nv_bfloat16* __restrict__ vx, long long ldkv, long long ldc,
                      const float* __restrict__ p_in, const __nv_bfloat16* __restrict__ dout, float scale,
                      __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dkc, __nv_bfloat16* __restrict__ dkx,
                      __nv_bfloat16* __restrict__ dvc, __nv_bfloat16* __restrict__ dvx, long long lddkv, long long lddc,
                      int B, int H, int n) {
    extern __shared__ __align__(16) float ca_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = n + 1, C = H * HD;
    float* sq = ca_smem + warp * (2 * HD + 2 * N);  // [HD] scaled q | [HD] dO | [N] p | [N] ds
    float* sdo = sq + HD;
    float* sp = sdo + HD;
    float* sds = sp + N;
    const int bh = blockIdx.x * CA_WARPS + warp;
    if (bh >= B * H) return;
    const int b = bh / H, h = bh - b * H;
    for (int e = lane; e < HD; e += 32) {
        sq[e] = __bfloat162float(q[(long long)b * C + h * HD + e]) * scale;
        sdo[e] = __bfloat162float(dout[(long long)b * C + h * HD + e]);
    }
    for (int j = lane; j < N; j += 32) sp[j] = p_in[(long long)bh * N + j];
    __syncwarp();
    // dp_j = dO . v_j ;  r = sum_j p_j dp_j
    float r = 0.f;
    for (int j = lane; j < N; j += 32) {
        const __nv_bfloat16* vr = (j == 0) ? vc + (long long)b * ldc + h * HD : vx + ((long long)b * n + j - 1) * ldkv + h * HD;
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
            const uint4 t = *reinterpret_cast<const uint4*>(vr + u * 8);
            a += sdo[u * 8 + 0] * bf16_lo(t.x) + sdo[u * 8 + 1] * bf16_hi(t.x) + sdo[u * 8 + 2] * bf16_lo(t.y) +
                 sdo[u * 8 + 3] * bf16_hi(t.y) + sdo[u * 8 + 4] * bf16_lo(t.z) + sdo[u * 8 + 5] * bf16_hi(t.z) +
                 sdo[u * 8 + 6] * bf16_lo(t.w) + sdo[u * 8 + 7] * bf16_hi(t.w);
        }
        sds[j] = a;
        r = fmaf(sp[j], a, r);
    }
    r = warp_sum(r);
    // ds_j = p_j (dp_j - r); dk_j = ds_j * (scale q); dv_j = p_j * dO   (each lane writes whole rows)
    for (int j = lane; j < N; j += 32) {
        const float ds = sp[j] * (sds[j] - r);
        sds[j] = ds;
        __nv_bfloat16* dkr = (j == 0) ? dkc + (long long)b * lddc + h * HD : dkx + ((long long)b * n + j - 1) * lddkv + h * HD;
        __nv_bfloat16* dvr = (j == 0) ? dvc + (long long)b * lddc + h * HD : dvx + ((long long)b * n + j - 1) * lddkv + h * HD;
        const float pj = sp[j];
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
            st_v4(dkr + u * 8, make_uint4(pack_bf16(ds * sq[u * 8 + 0], ds * sq[u * 8 + 1]),
                                          pack_bf16(ds * sq[u * 8 + 2], ds * sq[u * 8 + 3]),
                                          pack_bf16(ds * sq[u * 8 + 4], ds * sq[u * 8 + 5]),
                                          pack_bf16(ds * sq[u * 8 + 6], ds * sq[u * 8 + 7])));
            st_v4(dvr + u * 8, make_uint4(pack_bf16(pj * sdo[u * 8 + 0], pj * sdo[u * 8 + 1]),
                                          pack_bf16(pj * sdo[u * 8 + 2], pj * sdo[u * 8 + 3]),
                                          pack_bf16(pj * sdo[u * 8 + 4], pj * sdo[u * 8 + 5]),
                                          pack_bf16(pj * sdo[u * 8 + 6], pj * sdo[u * 8 + 7])));
        }
    }
    __syncwarp();
    // dq[e] = scale * sum_j ds_j k[j][e]
    const int e0 = lane * 2;
    if (e0 < HD) {
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < N; ++j) {
            const __nv_bfloat16* kr = (j == 0) ? kc + (long long)b * ldc + h * HD : kx + ((long long)b * n + j - 1) * ldkv + h * HD;
            const uint32_t t = *reinterpret_cast<const uint32_t*>(kr + e0);
            a0 = fmaf(sds[j], bf16_lo(t), a0);
            a1 = fmaf(sds[j], bf16_hi(t), a1);
        }
        *reinterpret_cast<uint32_t*>(dq + (long long)b * C + h * HD + e0) = pack_bf16(a0 * scale, a1 * scale);
    }
}

template <int H> static int th_fwd_launch(const float* S, const float* wl, const float* bl, const float* ww,
                                          const float* bw, float scale, __nv_bfloat16* Pm, float* rmax, float* rsum,
                                          int B, int N, int Np, cudaStream_t st) {
    const unsigned grid = (unsigned)((long long)B * N);
    if (Np <= 2 * TH_THREADS) th_mix_fwd_kernel<H, 2><<<grid, TH_THREADS, 0, st>>>(S, wl, bl, ww, bw, scale, Pm, rmax, rsum, B, N, Np);
    else if (Np <= 5 * TH_THREADS) th_mix_fwd_kernel<H, 5><<<grid, TH_THREADS, 0, st>>>(S, wl, bl, ww, bw, scale, Pm, rmax, rsum, B, N, Np);
    else if (Np <= 8 * TH_THREADS) th_mix_fwd_kernel<H, 8><<<grid, TH_THREADS, 0, st>>>(S, wl, bl, ww, bw, scale, Pm, rmax, rsum, B, N, Np);
    else return VITK_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

template <int H, int JT> static int th_bwd_launch_jt(const float* S, const float* dPm, const float* rmax,
                                                     const float* rsum, const float* wl, const float* bl,
                                                     const float* ww, const float* bw, float scale, __nv_bfloat16* dS,
                                                     float* dwl, float* dbl, float* dww, float* dbw, int B, int N,
                                                     int Np, cudaStream_t st) {
    const int smem = 4 * H * (Np + 1) * (int)sizeof(float);
    if (smem > 200 * 1024) return VITK_ERR_UNSUPPORTED;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(th_mix_bwd_kernel<H, JT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return VITK_ERR_CUDA;
        attr_smem = smem;
    }
    long long rows = (long long)B * N;
    long long per_sm = (200 * 1024) / (smem + 2048);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > rows) grid = rows;
    th_mix_bwd_kernel<H, JT><<<(unsigned)grid, TH_THREADS, smem, st>>>(S, dPm, rmax, rsum, wl, bl, ww, bw, scale, dS, dwl,
                                                                      dbl, dww, dbw, B, N, Np);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

template <int H> static int th_bwd_launch(const float* S, const float* dPm, const float* rmax, const float* rsum,
                                          const float* wl, const float* bl, const float* ww, const float* bw,
                                          float scale, __nv_bfloat16* dS, float* dwl, float* dbl, float* dww, float* dbw,
                                          int B, int N, int Np, cudaStream_t st) {
    if (Np <= 2 * TH_THREADS) return th_bwd_launch_jt<H, 2>(S, dPm, rmax, rsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
    if (Np <= 5 * TH_THREADS) return th_bwd_launch_jt<H, 5>(S, dPm, rmax, rsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
    if (Np <= 8 * TH_THREADS) return th_bwd_launch_jt<H, 8>(S, dPm, rmax, rsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
    return VITK_ERR_UNSUPPORTED;
}

}  // namespace vitk

using namespace vitk;

// version 2 (warp per row, tensor-core head mixes) takes rows of up to 16 * TH2_MAX_TILES keys; VITK_TH_MIX=1 forces v1
static bool th2_enabled(int Np) {
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("VITK_TH_MIX"); v1 = (e != nullptr && e[0] == '1') ? 1 : 0; }
    return !v1 && Np <= 16 * TH2_MAX_TILES;
}
template <int H, typename ST> static int th2_fwd_launch(const ST* S, const float* wl, const float* bl, const float* ww,
                                           const float* bw, float scale, __nv_bfloat16* Pm, float* rmax, float* rsum,
                                           int B, int N, int Np, cudaStream_t st) {
    const long long rows = (long long)B * N;
    long long grid = (rows + TH2_WARPS - 1) / TH2_WARPS;
    const long long cap = (long long)sm_count() * 4;
    if (grid > cap) grid = cap;
    if (Np <= 64) th_mix2_fwd_kernel<H, 4, ST><<<(unsigned)grid, TH2_WARPS * 32, 0, st>>>(S, wl, bl, ww, bw, scale, Pm, rmax, rsum, B, N, Np);
    else th_mix2_fwd_kernel<H, TH2_MAX_TILES, ST><<<(unsigned)grid, TH2_WARPS * 32, 0, st>>>(S, wl, bl, ww, bw, scale, Pm, rmax, rsum, B, N, Np);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
// VITK_TH_BWD = 0: original form (8 warps, 1 block / SM) | 1: lean, 8 warps x 2 blocks | 2: lean, 12 warps x 1 block (default for
// bf16 planes). fp32 planes always run the original form.
static int th2_bwd_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VITK_TH_BWD"); v = e ? atoi(e) : 2; if (v < 0 || v > 2) v = 2; }
    return v;
}
template <int H, typename ST, int WARPS, int MINB, bool LEAN>
static int th2_bwd_launch_v(const ST* S, const __nv_bfloat16* dPm, const float* rmax, const float* rsum, const float* wl,
                            const float* bl, const float* ww, float scale, __nv_bfloat16* dS, float* dwl, float* dbl,
                            float* dww, float* dbw, int B, int N, int Np, cudaStream_t st) {
    const long long rows = (long long)B * N;
    long long grid = (rows + WARPS - 1) / WARPS;
    const long long cap = (long long)sm_count() * MINB;      // persistent over rows
    if (grid > cap) grid = cap;
    if (Np <= 64)
        th_mix2_bwd_kernel<H, 4, ST, WARPS, MINB, LEAN><<<(unsigned)grid, WARPS * 32, 0, st>>>(
            S, dPm, rmax, rsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np);
    else
        th_mix2_bwd_kernel<H, TH2_MAX_TILES, ST, WARPS, MINB, LEAN><<<(unsigned)grid, WARPS * 32, 0, st>>>(
            S, dPm, rmax, rsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
template <int H, typename ST> static int th2_bwd_launch(const ST* S, const __nv_bfloat16* dPm, const float* rmax, const float* rsum,
                                                        const float* wl, const float* bl, const float* ww, float scale,
                                                        __nv_bfloat16* dS, float* dwl, float* dbl, float* dww, float* dbw,
                                                        int B, int N, int Np, cudaStream_t st) {
    if constexpr (sizeof(ST) == 2) {
        const int v = th2_bwd_variant();
        if (v == 1) return th2_bwd_launch_v<H, ST, 8, 2, true>(S, dPm, rmax, rsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        if constexpr (H <= 8) {     // (16 heads: the per-warp transposition scratch of 12 warps exceeds 48 KB)
            if (v == 2) return th2_bwd_launch_v<H, ST, 12, 1, true>(S, dPm, rmax, rsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        }
    }
    return th2_bwd_launch_v<H, ST, 8, 1, false>(S, dPm, rmax, rsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
}

extern "C" int vitk_th_mix_supports_bf16_dp(int Np) { return th2_enabled(Np) ? 1 : 0; }

template <typename ST>
static int th2_bwd_dispatch(const ST* S, const __nv_bfloat16* dP, const float* rowmax, const float* rowsum, const float* wl,
                            const float* bl, const float* ww, float scale, __nv_bfloat16* dS, float* dwl, float* dbl,
                            float* dww, float* dbw, int B, int H, int N, int Np, cudaStream_t st) {
    switch (H) {
        case 2: return th2_bwd_launch<2, ST>(S, dP, rowmax, rowsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 4: return th2_bwd_launch<4, ST>(S, dP, rowmax, rowsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 6: return th2_bwd_launch<6, ST>(S, dP, rowmax, rowsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 8: return th2_bwd_launch<8, ST>(S, dP, rowmax, rowsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 16: return th2_bwd_launch<16, ST>(S, dP, rowmax, rowsum, wl, bl, ww, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        default: return VITK_ERR_UNSUPPORTED;
    }
}
template <typename ST>
static int th2_fwd_dispatch(const ST* S, const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                            __nv_bfloat16* P, float* rowmax, float* rowsum, int B, int H, int N, int Np, cudaStream_t st) {
    switch (H) {
        case 2: return th2_fwd_launch<2, ST>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 4: return th2_fwd_launch<4, ST>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 6: return th2_fwd_launch<6, ST>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 8: return th2_fwd_launch<8, ST>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 16: return th2_fwd_launch<16, ST>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        default: return VITK_ERR_UNSUPPORTED;
    }
}

extern "C" int vitk_th_mix_bwd_bf16(const float* S, const void* dPm_bf16, const float* rowmax, const float* rowsum,
                                    const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                                    void* dS_bf16, float* dwl, float* dbl, float* dww, float* dbw, int B, int H, int N,
                                    int Np, void* stream) {
    if (B <= 0 || N <= 0 || Np < N || (Np % 8) != 0 || !S || !dPm_bf16 || !rowmax || !rowsum || !wl || !bl || !ww || !bw ||
        !dS_bf16 || !dwl || !dbl || !dww || !dbw)
        return VITK_ERR_ARG;
    if (!th2_enabled(Np)) return VITK_ERR_UNSUPPORTED;
    return th2_bwd_dispatch<float>(S, reinterpret_cast<const __nv_bfloat16*>(dPm_bf16), rowmax, rowsum, wl, bl, ww, scale,
                                   reinterpret_cast<__nv_bfloat16*>(dS_bf16), dwl, dbl, dww, dbw, B, H, N, Np,
                                   reinterpret_cast<cudaStream_t>(stream));
}

// bf16 logit planes (what the reference's matmul produces under torch.autocast(bfloat16)): version-2 kernels only
extern "C" int vitk_th_mix_fwd_s16(const void* S_bf16, const float* wl, const float* bl, const float* ww, const float* bw,
                                   float scale, void* Pm_bf16, float* rowmax, float* rowsum, int B, int H, int N, int Np,
                                   void* stream) {
    if (B <= 0 || N <= 0 || Np < N || (Np % 8) != 0 || !S_bf16 || !wl || !bl || !ww || !bw || !Pm_bf16 || !rowmax || !rowsum)
        return VITK_ERR_ARG;
    if (!th2_enabled(Np)) return VITK_ERR_UNSUPPORTED;
    return th2_fwd_dispatch<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(S_bf16), wl, bl, ww, bw, scale,
                                           reinterpret_cast<__nv_bfloat16*>(Pm_bf16), rowmax, rowsum, B, H, N, Np,
                                           reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vitk_th_mix_bwd_s16(const void* S_bf16, const void* dPm_bf16, const float* rowmax, const float* rowsum,
                                   const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                                   void* dS_bf16, float* dwl, float* dbl, float* dww, float* dbw, int B, int H, int N,
                                   int Np, void* stream) {
    if (B <= 0 || N <= 0 || Np < N || (Np % 8) != 0 || !S_bf16 || !dPm_bf16 || !rowmax || !rowsum || !wl || !bl || !ww ||
        !bw || !dS_bf16 || !dwl || !dbl || !dww || !dbw)
        return VITK_ERR_ARG;
    if (!th2_enabled(Np)) return VITK_ERR_UNSUPPORTED;
    return th2_bwd_dispatch<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(S_bf16),
                                           reinterpret_cast<const __nv_bfloat16*>(dPm_bf16), rowmax, rowsum, wl, bl, ww,
                                           scale, reinterpret_cast<__nv_bfloat16*>(dS_bf16), dwl, dbl, dww, dbw, B, H, N, Np,
                                           reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vitk_th_mix_fwd(const float* S, const float* wl, const float* bl, const float* ww, const float* bw,
                               float scale, void* Pm_bf16, float* rowmax, float* rowsum, int B, int H, int N, int Np,
                               void* stream) {
    if (B <= 0 || N <= 0 || Np < N || (Np % 8) != 0 || !S || !wl || !bl || !ww || !bw || !Pm_bf16 || !rowmax || !rowsum)
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto P = reinterpret_cast<__nv_bfloat16*>(Pm_bf16);
    if (th2_enabled(Np)) return th2_fwd_dispatch<float>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, H, N, Np, st);
    switch (H) {
        case 4: return th_fwd_launch<4>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 6: return th_fwd_launch<6>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 8: return th_fwd_launch<8>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 16: return th_fwd_launch<16>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        case 2: return th_fwd_launch<2>(S, wl, bl, ww, bw, scale, P, rowmax, rowsum, B, N, Np, st);
        default: return VITK_ERR_UNSUPPORTED;
    }
}

extern "C" int vitk_th_mix_bwd(const float* S, const float* dPm, const float* rowmax, const float* rowsum,
                               const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                               void* dS_bf16, float* dwl, float* dbl, float* dww, float* dbw, int B, int H, int N,
                               int Np, void* stream) {
    if (B <= 0 || N <= 0 || Np < N || (Np % 8) != 0 || !S || !dPm || !rowmax || !rowsum || !wl || !bl || !ww || !bw ||
        !dS_bf16 || !dwl || !dbl || !dww || !dbw)
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto dS = reinterpret_cast<__nv_bfloat16*>(dS_bf16);
    switch (H) {
        case 4: return th_bwd_launch<4>(S, dPm, rowmax, rowsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 6: return th_bwd_launch<6>(S, dPm, rowmax, rowsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 8: return th_bwd_launch<8>(S, dPm, rowmax, rowsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 16: return th_bwd_launch<16>(S, dPm, rowmax, rowsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        case 2: return th_bwd_launch<2>(S, dPm, rowmax, rowsum, wl, bl, ww, bw, scale, dS, dwl, dbl, dww, dbw, B, N, Np, st);
        default: return VITK_ERR_UNSUPPORTED;
    }
}

extern "C" int vitk_class_attn_fwd(const void* q, const void* kc, const void* kx, const void* vc, const void* vx,
                                   long long ldkv, long long ldc, float scale, void* out, float* p, int B, int H, int n,
                                   int d, void* stream) {
    if (B <= 0 || H <= 0 || n < 0 || !(d == 48 || d == 64) || (ldkv % 8) != 0 || (ldc % 8) != 0 || !q || !kc || !vc || !out || !p ||
        (n > 0 && (!kx || !vx)))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int N = n + 1;
    const int smem = CA_WARPS * (d + N) * (int)sizeof(float);
    if (smem > 48 * 1024) return VITK_ERR_UNSUPPORTED;
    const int grid = (B * H + CA_WARPS - 1) / CA_WARPS;
    auto b = [](const void* x) { return reinterpret_cast<const __nv_bfloat16*>(x); };
    if (d == 48)
        class_attn_fwd_kernel<48><<<grid, CA_WARPS * 32, smem, st>>>(b(q), b(kc), b(kx), b(vc), b(vx), ldkv, ldc, scale,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), p, B, H, n);
    else
        class_attn_fwd_kernel<64><<<grid, CA_WARPS * 32, smem, st>>>(b(q), b(kc), b(kx), b(vc), b(vx), ldkv, ldc, scale,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), p, B, H, n);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_class_attn_bwd(const void* q, const void* kc, const void* kx, const void* vc, const void* vx,
                                   long long ldkv, long long ldc, const float* p, const void* dout, float scale,
                                   void* dq, void* dkc, void* dkx, void* dvc, void* dvx, long long lddkv, long long lddc,
                                   int B, int H, int n, int d, void* stream) {
    if (B <= 0 || H <= 0 || n < 0 || !(d == 48 || d == 64) || ((ldkv | lddkv | ldc | lddc) % 8) != 0 || !q || !kc || !vc ||
        !p || !dout || !dq || !dkc || !dvc || (n > 0 && (!kx || !vx || !dkx || !dvx)))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int N = n + 1;
    const int smem = CA_WARPS * (2 * d + 2 * N) * (int)sizeof(float);
    if (smem > 48 * 1024) return VITK_ERR_UNSUPPORTED;
    const int grid = (B * H + CA_WARPS - 1) / CA_WARPS;
    auto b = [](const void* x) { return reinterpret_cast<const __nv_bfloat16*>(x); };
    auto m = [](void* x) { return reinterpret_cast<__nv_bfloat16*>(x); };
    if (d == 48)
        class_attn_bwd_kernel<48><<<grid, CA_WARPS * 32, smem, st>>>(b(q), b(kc), b(kx), b(vc), b(vx), ldkv, ldc, p, b(dout),
                                                                    scale, m(dq), m(dkc), m(dkx), m(dvc), m(dvx), lddkv,
                                                                    lddc, B, H, n);
    else
        class_attn_bwd_kernel<64><<<grid, CA_WARPS * 32, smem, st>>>(b(q), b(kc), b(kx), b(vc), b(vx), ldkv, ldc, p, b(dout),
                                                                    scale, m(dq), m(dkc), m(dkx), m(dvc), m(dvx), lddkv,
                                                                    lddc, B, H, n);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}
