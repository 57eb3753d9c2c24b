// Memory-bound glue kernels of the ViT block: LayerNorm fwd/bwd (one warp per row, 128-bit loads, shuffle
// reductions), bias-gradient column sums, dtype casts. All HBM-roofline kernels: no smem staging needed because
// every element is touched once; grids are multiples of the SM count with grid-stride loops.
#include "common.cuh"
#include <stdlib.h>
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int LN_WARPS = 8;

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: x fp32 [rows, D] -> y bf16, mean/rstd fp32.    (nn.LayerNorm, eps inside sqrt, biased variance)
// ------------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ w,
              const float* __restrict__ b, __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 wv[MAXC], bv[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = (i * 32 + lane) * 4;
        wv[i] = (c < D) ? __ldg(reinterpret_cast<const float4*>(w + c)) : make_float4(0, 0, 0, 0);
        bv[i] = (c < D) ? __ldg(reinterpret_cast<const float4*>(b + c)) : make_float4(0, 0, 0, 0);
    }
    const float inv_d = 1.0f / static_cast<float>(D);
    for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
        const float* xr = x + row * x_stride;
        float4 v[MAXC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            v[i] = (c < D) ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0, 0, 0, 0);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
                q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float o0 = (v[i].x - mean) * rstd * wv[i].x + bv[i].x;
                const float o1 = (v[i].y - mean) * rstd * wv[i].y + bv[i].y;
                const float o2 = (v[i].z - mean) * rstd * wv[i].z + bv[i].z;
                const float o3 = (v[i].w - mean) * rstd * wv[i].w + bv[i].w;
                if (y != nullptr) *reinterpret_cast<uint2*>(y + row * D + c) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
                if (y32 != nullptr) *reinterpret_cast<float4*>(y32 + row * D + c) = make_float4(o0, o1, o2, o3);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward fused with residual-gradient add and the bf16 copy the next GEMM consumes.
//   dx = dres + rstd * (dy*w - mean_D(dy*w) - xhat * mean_D(dy*w*xhat));  dw += sum_rows dy*xhat;  db += sum_rows dy
// ------------------------------------------------------------------------------------------------
template <int MAXC, bool DY_F32>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, long long x_stride,
              const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ dres, float* __restrict__ dx, long long dx_stride,
              __nv_bfloat16* __restrict__ dx_bf16, const float* __restrict__ colscale, float* __restrict__ dweight,
              float* __restrict__ dbias, float* __restrict__ dxsum, long long rows, int D) {
    // smem: w[Dp] | per-warp partials [LN_WARPS][3][Dp] (dweight, dbias, column sums of the bf16 dx copy). Keeping the partial sums and the weight in
    // shared memory instead of registers leaves room for two 8-warp blocks per SM with all loads of a row in flight.
    constexpr int Dp = MAXC * 128;
    extern __shared__ __align__(16) float ln_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sw = ln_smem;
    float* pw = ln_smem + Dp + warp * 3 * Dp;
    float* pb = pw + Dp;
    float* px = pb + Dp;
    for (int c = threadIdx.x; c < Dp; c += LN_WARPS * 32) sw[c] = (c < D) ? w[c] : 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = (i * 32 + lane) * 4;
        *reinterpret_cast<float4*>(pw + c) = make_float4(0, 0, 0, 0);
        *reinterpret_cast<float4*>(pb + c) = make_float4(0, 0, 0, 0);
        *reinterpret_cast<float4*>(px + c) = make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    const float inv_d = 1.0f / static_cast<float>(D);
    for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
        const float* xr = x + row * x_stride;
        float4 xv[MAXC], dv[MAXC], rv[MAXC];
        // issue every global load of the row before the first use
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                xv[i] = *reinterpret_cast<const float4*>(xr + c);
                if constexpr (DY_F32) {
                    dv[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * D + c);
                } else {
                    const uint2 t =
                        *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + row * D + c);
                    dv[i] = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
                }
                rv[i] = (dres != nullptr) ? *reinterpret_cast<const float4*>(dres + row * dx_stride + c)
                                          : make_float4(0, 0, 0, 0);
            } else {
                xv[i] = make_float4(0, 0, 0, 0);
                dv[i] = make_float4(0, 0, 0, 0);
                rv[i] = make_float4(0, 0, 0, 0);
            }
        }
        const float mean = mean_in[row], rstd = rstd_in[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float4 wv = *reinterpret_cast<const float4*>(sw + c);
                // xv <- xhat ; dv stays dy ; g = dy * w
                xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                                    (xv[i].w - mean) * rstd);
                float4 aw = *reinterpret_cast<float4*>(pw + c);
                float4 ab = *reinterpret_cast<float4*>(pb + c);
                aw.x += dv[i].x * xv[i].x; aw.y += dv[i].y * xv[i].y; aw.z += dv[i].z * xv[i].z; aw.w += dv[i].w * xv[i].w;
                ab.x += dv[i].x; ab.y += dv[i].y; ab.z += dv[i].z; ab.w += dv[i].w;
                *reinterpret_cast<float4*>(pw + c) = aw;
                *reinterpret_cast<float4*>(pb + c) = ab;
                dv[i] = make_float4(dv[i].x * wv.x, dv[i].y * wv.y, dv[i].z * wv.z, dv[i].w * wv.w);
                s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
                s2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
            }
        }
        const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float4 o = make_float4(rstd * (dv[i].x - c1 - xv[i].x * c2) + rv[i].x,
                                       rstd * (dv[i].y - c1 - xv[i].y * c2) + rv[i].y,
                                       rstd * (dv[i].z - c1 - xv[i].z * c2) + rv[i].z,
                                       rstd * (dv[i].w - c1 - xv[i].w * c2) + rv[i].w);
                if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * dx_stride + c) = o;
                if (dx_bf16 != nullptr) {
                    if (colscale != nullptr) {
                        const float4 cs = __ldg(reinterpret_cast<const float4*>(colscale + c));
                        o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
                    }
                    *reinterpret_cast<uint2*>(dx_bf16 + row * D + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                    if (dxsum != nullptr) {  // bias gradient of the Linear that consumes the bf16 copy
                        float4 ax = *reinterpret_cast<float4*>(px + c);
                        ax.x += o.x; ax.y += o.y; ax.z += o.z; ax.w += o.w;
                        *reinterpret_cast<float4*>(px + c) = ax;
                    }
                }
            }
        }
    }
    // block reduction of the per-warp partials, then one atomic per column per block
    __syncthreads();
    const float* part = ln_smem + Dp;
    for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
        float a = 0.f, bsum = 0.f, xsum = 0.f;
#pragma unroll
        for (int k = 0; k < LN_WARPS; ++k) {
            a += part[k * 3 * Dp + c];
            bsum += part[k * 3 * Dp + Dp + c];
            xsum += part[k * 3 * Dp + 2 * Dp + c];
        }
        if (dweight != nullptr) atomicAdd(dweight + c, a);
        if (dbias != nullptr) atomicAdd(dbias + c, bsum);
        if (dxsum != nullptr) atomicAdd(dxsum + c, xsum);
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, TMA-staged variant for large row counts. Same math and outputs as ln_bwd_kernel.
// ncu on ln_bwd_kernel (25216 x 768): 4.08 TB/s = 62 % of HBM peak with 70 % of the samples on the long scoreboard --
// a warp cannot have the next row's loads in flight while it reduces the current one. Here a producer thread streams
// whole rows (x fp32 | dres fp32 | dy) into a shared-memory ring with 1-D bulk async copies (cp.async.bulk, completion
// counted on an mbarrier per stage), so LNR_STAGES rows per SM are always in flight; eight consumer warps take the
// rows round-robin, keep xhat / dy*w of the row and the dweight / dbias / column-sum partials in registers, and store
// dx (fp32) and its bf16 copy straight from registers.
// A stage holds R consecutive rows (R > 1 when the rows are contiguous in memory): the single producer thread issues
// three bulk copies per STAGE, and at D = 384 one row per stage left the kernel bound by that thread's issue rate
// (50 us for 154 MB, 510 copies per SM); a block owns a contiguous range of rows for this.
// ------------------------------------------------------------------------------------------------
constexpr int LNR_CONSUMERS = 8;
constexpr int LNR_THREADS = (LNR_CONSUMERS + 1) * 32;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int MAXC, bool DY_F32>
__global__ void __launch_bounds__(LNR_THREADS, 1)
ln_bwd_ring_kernel(const void* __restrict__ dy_, const float* __restrict__ x, long long x_stride,
                   const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   const float* __restrict__ dres, float* __restrict__ dx, long long dx_stride,
                   __nv_bfloat16* __restrict__ dx_bf16, const float* __restrict__ colscale, float* __restrict__ dweight,
                   float* __restrict__ dbias, float* __restrict__ dxsum, long long rows, int D, int stages, int R) {
    constexpr int Dp = MAXC * 128;
    extern __shared__ __align__(128) uint8_t lnr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x_bytes = D * 4, r_bytes = (dres != nullptr) ? D * 4 : 0, dy_bytes = D * (DY_F32 ? 4 : 2);
    const uint32_t stage_bytes = ((uint32_t)R * (x_bytes + D * 4 + D * (DY_F32 ? 4 : 2)) + 127) & ~127u;
    uint8_t* ring = lnr_smem;
    float* sw = reinterpret_cast<float*>(lnr_smem + (size_t)stages * stage_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(sw + Dp);
    uint64_t* empty = full + stages;
    for (int c = threadIdx.x; c < Dp; c += LNR_THREADS) sw[c] = (c < D) ? w[c] : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], R);      // one arrival per row of the stage
        }
        fence_mbar_init();
    }
    __syncthreads();
    // rows of this block: the contiguous range [row0, row0 + nloc)
    const long long per_block = (rows + gridDim.x - 1) / gridDim.x;
    const long long row0 = (long long)blockIdx.x * per_block;
    const long long nloc = row0 >= rows ? 0 : (rows - row0 < per_block ? rows - row0 : per_block);
    const long long ngroups = (nloc + R - 1) / R;
    float4 aw[MAXC], ab[MAXC], ax[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) aw[i] = ab[i] = ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (warp == LNR_CONSUMERS) {
        // ===================== producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long g = 0; g < ngroups; ++g) {
                const long long row = row0 + g * R;
                const uint32_t nr = (uint32_t)(nloc - g * R < R ? nloc - g * R : R);    // rows of this stage
                // (a short last group never completes its `empty` phase: nothing waits on it any more)
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = ring + (size_t)s * stage_bytes;
                mbar_expect_tx(&full[s], nr * (x_bytes + r_bytes + dy_bytes));
                // (R > 1 only with x_stride == dx_stride == D: the rows of a stage are one contiguous range)
                bulk_g2s(st, x + row * x_stride, nr * x_bytes, &full[s]);
                if (r_bytes) bulk_g2s(st + (size_t)R * x_bytes, dres + row * dx_stride, nr * r_bytes, &full[s]);
                bulk_g2s(st + (size_t)2 * R * x_bytes, reinterpret_cast<const uint8_t*>(dy_) + (size_t)row * dy_bytes,
                         nr * dy_bytes, &full[s]);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== consumers =====================
        const float inv_d = 1.0f / static_cast<float>(D);
        for (long long k = warp; k < nloc; k += LNR_CONSUMERS) {
            const long long row = row0 + k;
            const long long g = k / R;
            const int rr = static_cast<int>(k - g * R);          // row inside its stage
            const int s = static_cast<int>(g % stages);
            const uint32_t ph = static_cast<uint32_t>((g / stages) & 1);
            const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
            mbar_wait(&full[s], ph);
            const uint8_t* st = ring + (size_t)s * stage_bytes;
            const float* sx = reinterpret_cast<const float*>(st + (size_t)rr * x_bytes);
            const float* sr = reinterpret_cast<const float*>(st + (size_t)(R + rr) * x_bytes);
            const uint8_t* sdy = st + (size_t)2 * R * x_bytes + (size_t)rr * dy_bytes;
            float4 xh[MAXC], gv[MAXC];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const int c = (i * 32 + lane) * 4;
                if (c < D) {
                    const float4 xv = *reinterpret_cast<const float4*>(sx + c);
                    float4 dv;
                    if constexpr (DY_F32) {
                        dv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sdy) + c);
                    } else {
                        const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(sdy) + c);
                        dv = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
                    }
                    const float4 wv = *reinterpret_cast<const float4*>(sw + c);
                    xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                                        (xv.w - mean) * rstd);
                    aw[i].x += dv.x * xh[i].x; aw[i].y += dv.y * xh[i].y; aw[i].z += dv.z * xh[i].z; aw[i].w += dv.w * xh[i].w;
                    ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
                    gv[i] = make_float4(dv.x * wv.x, dv.y * wv.y, dv.z * wv.z, dv.w * wv.w);
                    s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
                    s2 += (gv[i].x * xh[i].x + gv[i].y * xh[i].y) + (gv[i].z * xh[i].z + gv[i].w * xh[i].w);
                } else {
                    xh[i] = gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
            float4 rv[MAXC];
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const int c = (i * 32 + lane) * 4;
                rv[i] = (r_bytes && c < D) ? *reinterpret_cast<const float4*>(sr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // every shared-memory read of the stage is done: hand it back to the producer before the global stores
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const int c = (i * 32 + lane) * 4;
                if (c < D) {
                    float4 o = make_float4(rstd * (gv[i].x - c1 - xh[i].x * c2) + rv[i].x,
                                           rstd * (gv[i].y - c1 - xh[i].y * c2) + rv[i].y,
                                           rstd * (gv[i].z - c1 - xh[i].z * c2) + rv[i].z,
                                           rstd * (gv[i].w - c1 - xh[i].w * c2) + rv[i].w);
                    if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * dx_stride + c) = o;
                    if (dx_bf16 != nullptr) {
                        if (colscale != nullptr) {
                            const float4 cs = __ldg(reinterpret_cast<const float4*>(colscale + c));
                            o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
                        }
                        *reinterpret_cast<uint2*>(dx_bf16 + row * D + c) =
                            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                        ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
                    }
                }
            }
        }
    }
    // block reduction of the per-warp register partials through the (now idle) ring, one atomic per column per block
    __syncthreads();
    float* part = reinterpret_cast<float*>(ring);  // [LNR_CONSUMERS][3][Dp]
    if (warp < LNR_CONSUMERS) {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            *reinterpret_cast<float4*>(part + (warp * 3 + 0) * Dp + c) = aw[i];
            *reinterpret_cast<float4*>(part + (warp * 3 + 1) * Dp + c) = ab[i];
            *reinterpret_cast<float4*>(part + (warp * 3 + 2) * Dp + c) = ax[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LNR_THREADS) {
        float a = 0.f, bsum = 0.f, xsum = 0.f;
#pragma unroll
        for (int k = 0; k < LNR_CONSUMERS; ++k) {
            a += part[(k * 3 + 0) * Dp + c];
            bsum += part[(k * 3 + 1) * Dp + c];
            xsum += part[(k * 3 + 2) * Dp + c];
        }
        if (dweight != nullptr) atomicAdd(dweight + c, a);
        if (dbias != nullptr) atomicAdd(dbias + c, bsum);
        if (dxsum != nullptr) atomicAdd(dxsum + c, xsum);
    }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix (bias gradient): out[n] += sum_rows x[row, n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long rows, int N, float* __restrict__ out) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col + 8 <= N) {
        for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
            const uint4 v = ld_nc_v4(x + r * ldx + col);
            acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
            acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    const long long n8 = n / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = *reinterpret_cast<const float4*>(x + i * 8);
        const float4 b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
        st_v4(y + i * 8, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w)));
    }
    if (blockIdx.x == 0) {
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) y[i] = __float2bfloat16_rn(x[i]);
    }
}


// ------------------------------------------------------------------------------------------------
// Multi-tensor SGD with momentum (torch.optim.SGD semantics: dampening 0, no nesterov, no weight decay;
// reference optimiser utils_network.py:119-126) fused with the refresh of the bf16 GEMM-operand copy of the weight:
//   buf = momentum * buf + g * gscale ;  p -= lr * buf ;  w_bf16 = bf16(p)
// One launch for the whole model: `tab` holds per-tensor pointers, `chunks` maps each thread block to
// (tensor, first element). Reads 12 B, writes 8 (+2) B per parameter.
// ------------------------------------------------------------------------------------------------
struct SgdTensor {
    float* p;
    const float* g;
    float* buf;
    __nv_bfloat16* w;  // may be null
    long long n;
};
constexpr int SGD_CHUNK = 16384;  // elements per thread block

__global__ void __launch_bounds__(256)
sgd_momentum_multi_kernel(const SgdTensor* __restrict__ tab, const int2* __restrict__ chunks, float lr, float momentum,
                          float gscale, int first_step, const float* __restrict__ hp) {
    if (hp != nullptr) {   // hyper-parameters from device memory: a captured CUDA graph follows LambdaLR without re-capture
        lr = hp[0];
        momentum = hp[1];
        gscale = hp[2];
    }
    const int2 ck = chunks[blockIdx.x];  // x = tensor index, y = chunk index within the tensor
    const SgdTensor t = tab[ck.x];
    const long long base = (long long)ck.y * SGD_CHUNK;
    const long long end = min(base + SGD_CHUNK, t.n);
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                       reinterpret_cast<uintptr_t>(t.buf)) & 15) == 0 &&
                     (t.w == nullptr || (reinterpret_cast<uintptr_t>(t.w) & 7) == 0);
    if (vec) {
        const long long end4 = base + ((end - base) & ~3LL);
        for (long long i = base + threadIdx.x * 4; i < end4; i += 256 * 4) {
            const float4 g = *reinterpret_cast<const float4*>(t.g + i);
            float4 p = *reinterpret_cast<const float4*>(t.p + i);
            float4 b = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(t.buf + i);
            b.x = fmaf(momentum, b.x, g.x * gscale); b.y = fmaf(momentum, b.y, g.y * gscale);
            b.z = fmaf(momentum, b.z, g.z * gscale); b.w = fmaf(momentum, b.w, g.w * gscale);
            p.x = fmaf(-lr, b.x, p.x); p.y = fmaf(-lr, b.y, p.y); p.z = fmaf(-lr, b.z, p.z); p.w = fmaf(-lr, b.w, p.w);
            *reinterpret_cast<float4*>(t.buf + i) = b;
            *reinterpret_cast<float4*>(t.p + i) = p;
            if (t.w != nullptr) *reinterpret_cast<uint2*>(t.w + i) = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
        }
        for (long long i = end4 + threadIdx.x; i < end; i += 256) {
            const float b = fmaf(momentum, first_step ? 0.f : t.buf[i], t.g[i] * gscale);
            const float p = fmaf(-lr, b, t.p[i]);
            t.buf[i] = b; t.p[i] = p;
            if (t.w != nullptr) t.w[i] = __float2bfloat16_rn(p);
        }
    } else {
        for (long long i = base + threadIdx.x; i < end; i += 256) {
            const float b = fmaf(momentum, first_step ? 0.f : t.buf[i], t.g[i] * gscale);
            const float p = fmaf(-lr, b, t.p[i]);
            t.buf[i] = b; t.p[i] = p;
            if (t.w != nullptr) t.w[i] = __float2bfloat16_rn(p);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-tensor Adam / AdamW (torch.optim.Adam / AdamW semantics without amsgrad / maximize; the reference's 'adam' and
// 'adamw' entries, utils_network.py:119-126) fused with the bf16 weight refresh, same table / chunk scheme as SGD.
// All hyper-parameters AND the step counter live in device memory `hp` (so a captured CUDA graph advances the bias
// corrections by itself): hp = { lr, beta1, beta2, eps, weight_decay, decoupled (1 = AdamW), 1-beta1, 1-beta2 (rounded
// from double on the host, as torch does), step, step_size, inv_sqrt_bias2 }. adam_tick_kernel advances step and derives
// the last two in double precision, once per launch.
//   AdamW: p *= 1 - lr*wd           Adam: g += wd*p
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= step_size * m / (sqrt(v) * inv_sqrt_bias2 + eps)
// ------------------------------------------------------------------------------------------------
struct AdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    __nv_bfloat16* w;  // may be null
    long long n;
};

__global__ void adam_tick_kernel(float* __restrict__ hp) {
    const double step = (double)hp[8] + 1.0;
    const double b1 = hp[1], b2 = hp[2];
    const double bias1 = 1.0 - pow(b1, step), bias2 = 1.0 - pow(b2, step);
    hp[8] = (float)step;
    hp[9] = (float)((double)hp[0] / bias1);
    hp[10] = (float)(1.0 / sqrt(bias2));
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float lr, float b1, float b2,
                                            float omb1, float omb2, float eps, float wd, bool decoupled,
                                            float step_size, float isb2) {
    if (wd != 0.f) {
        if (decoupled) p *= 1.f - lr * wd;
        else g = fmaf(wd, p, g);
    }
    m = fmaf(b1, m, omb1 * g);             // torch: exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(b2, v, omb2 * g * g);
    const float denom = fmaf(sqrtf(v), isb2, eps);
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks, const float* __restrict__ hp) {
    const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], omb1 = hp[6], omb2 = hp[7];
    const float step_size = hp[9], isb2 = hp[10];
    const bool decoupled = hp[5] != 0.f;
    const int2 ck = chunks[blockIdx.x];
    const AdamTensor t = tab[ck.x];
    const long long base = (long long)ck.y * SGD_CHUNK;
    const long long end = min(base + SGD_CHUNK, t.n);
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                       reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                     (t.w == nullptr || (reinterpret_cast<uintptr_t>(t.w) & 7) == 0);
    const long long end4 = vec ? base + ((end - base) & ~3LL) : base;
    for (long long i = base + threadIdx.x * 4; i < end4; i += 256 * 4) {
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 p = *reinterpret_cast<const float4*>(t.p + i);
        float4 m = *reinterpret_cast<const float4*>(t.m + i);
        float4 v = *reinterpret_cast<const float4*>(t.v + i);
        adam_update(p.x, g.x, m.x, v.x, lr, b1, b2, omb1, omb2, eps, wd, decoupled, step_size, isb2);
        adam_update(p.y, g.y, m.y, v.y, lr, b1, b2, omb1, omb2, eps, wd, decoupled, step_size, isb2);
        adam_update(p.z, g.z, m.z, v.z, lr, b1, b2, omb1, omb2, eps, wd, decoupled, step_size, isb2);
        adam_update(p.w, g.w, m.w, v.w, lr, b1, b2, omb1, omb2, eps, wd, decoupled, step_size, isb2);
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
        *reinterpret_cast<float4*>(t.p + i) = p;
        if (t.w != nullptr) *reinterpret_cast<uint2*>(t.w + i) = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
    }
    for (long long i = end4 + threadIdx.x; i < end; i += 256) {
        float p = t.p[i], m = t.m[i], v = t.v[i];
        adam_update(p, t.g[i], m, v, lr, b1, b2, omb1, omb2, eps, wd, decoupled, step_size, isb2);
        t.m[i] = m; t.v[i] = v; t.p[i] = p;
        if (t.w != nullptr) t.w[i] = __float2bfloat16_rn(p);
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 column sums with a row stride: out[c] += sum_r x[r*ldx + c]   (d_pos = sum_b dX[b], d_cls)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long ldx, long long rows, long long cols, float* __restrict__ out) {
    const long long c = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= cols) return;
    float4 acc = make_float4(0, 0, 0, 0);
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (gridDim.y == 1) {
        float4 o = *reinterpret_cast<float4*>(out + c);
        o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
        *reinterpret_cast<float4*>(out + c) = o;
    } else {
        red_add_v4_f32(out + c, acc.x, acc.y, acc.z, acc.w);
    }
}

// ------------------------------------------------------------------------------------------------
// out[c] += sum_r a_f32[r, c] * b_bf16[r, c]     (LayerScale: d_gamma = sum_rows dY o f, models/cait.py:144-149)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_prod_kernel(const float* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b, long long ldb,
                   long long rows, int N, float* __restrict__ out, const float* __restrict__ rowscale, long long rps) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col + 8 <= N) {
        for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
            const uint4 v = ld_nc_v4(b + r * ldb + col);
            float4 a0 = *reinterpret_cast<const float4*>(a + r * lda + col);
            float4 a1 = *reinterpret_cast<const float4*>(a + r * lda + col + 4);
            if (rowscale != nullptr) {      // DropPath: the branch was scaled by mask / keep_prob per sample
                const float rs = __ldg(rowscale + r / rps);
                a0.x *= rs; a0.y *= rs; a0.z *= rs; a0.w *= rs;
                a1.x *= rs; a1.y *= rs; a1.z *= rs; a1.w *= rs;
            }
            acc[0] += a0.x * bf16_lo(v.x); acc[1] += a0.y * bf16_hi(v.x); acc[2] += a0.z * bf16_lo(v.y);
            acc[3] += a0.w * bf16_hi(v.y); acc[4] += a1.x * bf16_lo(v.z); acc[5] += a1.y * bf16_hi(v.z);
            acc[6] += a1.z * bf16_lo(v.w); acc[7] += a1.w * bf16_hi(v.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

// ------------------------------------------------------------------------------------------------
// LayerScale branch backward in ONE pass (models/cait.py:144-149: x + gamma * DropPath(f)):
//   out_bf16[r, c] = bf16(dy[r, c] * gamma[c] * rs[r])     the gradient of the branch output f, A operand of the
//                                                           proj / fc2 dgrad and wgrad GEMMs
//   dgamma[c]  += sum_r dy[r, c] * rs[r] * f[r, c]
//   dbias[c]   += sum_r dy[r, c] * gamma[c] * rs[r]         the bias gradient of the Linear that produced f
// (replaces colsum_prod + scale_cast + colsum_bf16: three passes over [rows, D] per branch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layerscale_bwd_kernel(const float* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ f, long long ldf,
                      const float* __restrict__ gamma, const float* __restrict__ rowscale, long long rps, long long rows,
                      int N, __nv_bfloat16* __restrict__ out, float* __restrict__ dgamma, float* __restrict__ dbias) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col + 8 <= N) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
            const uint4 v = ld_nc_v4(f + r * ldf + col);
            const float4 a0 = *reinterpret_cast<const float4*>(dy + r * lddy + col);
            const float4 a1 = *reinterpret_cast<const float4*>(dy + r * lddy + col + 4);
            const float rs = rowscale != nullptr ? __ldg(rowscale + r / rps) : 1.0f;
            const float a[8] = {a0.x * rs, a0.y * rs, a0.z * rs, a0.w * rs, a1.x * rs, a1.y * rs, a1.z * rs, a1.w * rs};
            const float fv[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                                 bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ag[j] = fmaf(a[j], fv[j], ag[j]);
                o[j] = a[j] * gm[j];
                ab[j] += o[j];
            }
            st_v4(out + r * N + col, make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]),
                                                pack_bf16(o[6], o[7])));
        }
    }
    const int c = blockIdx.x * 256 + threadIdx.x;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dgamma : dbias;
        if (dst == nullptr) continue;       // (uniform)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? ag[j] : ab[j];
        __syncthreads();
        if (c < N) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += red[k][threadIdx.x];
            atomicAdd(dst + c, sum);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 with optional per-column / per-sample scale and token-row compaction:
//   out[r, :] = bf16(x[(r / rpg) * group_stride + (r % rpg) * D + :] * colscale[:] * rowscale[r / rps])
// (bf16 copy of the residual gradient the dgrad/wgrad GEMMs consume; dX[:, T:, :] -> patch rows for PatchEmbed wgrad)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scale_cast_kernel(const float* __restrict__ x, long long rpg, long long group_stride, long long rows, int D,
                  const float* __restrict__ colscale, const float* __restrict__ rowscale, long long rps,
                  __nv_bfloat16* __restrict__ out) {
    const int vec_per_row = D / 8;
    const long long total = rows * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec_per_row;
        const int c = static_cast<int>(i - r * vec_per_row) * 8;
        const float* src = x + (r / rpg) * group_stride + (r % rpg) * D + c;
        float4 a = *reinterpret_cast<const float4*>(src);
        float4 b = *reinterpret_cast<const float4*>(src + 4);
        if (colscale != nullptr) {
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(colscale + c));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(colscale + c + 4));
            a.x *= c0.x; a.y *= c0.y; a.z *= c0.z; a.w *= c0.w;
            b.x *= c1.x; b.y *= c1.y; b.z *= c1.z; b.w *= c1.w;
        }
        if (rowscale != nullptr) {
            const float rs = __ldg(rowscale + r / rps);
            a.x *= rs; a.y *= rs; a.z *= rs; a.w *= rs;
            b.x *= rs; b.y *= rs; b.z *= rs; b.w *= rs;
        }
        st_v4(out + r * D + c, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w)));
    }
}

// ------------------------------------------------------------------------------------------------
// patchify (v1 of PatchEmbed's A operand): x fp32 [B,C,H,W] -> bf16 [B*n, C*P*P], k = (c, i, j)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int P) {
    const int gw = W / P, gh = H / P;
    const int K = C * P * P;
    const int vec_per_row = K / 4;
    const long long total = (long long)B * gh * gw * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec_per_row;
        const int k = static_cast<int>(i - r * vec_per_row) * 4;
        const int c = k / (P * P), ij = k - c * P * P, ii = ij / P, jj = ij - ii * P;
        const int b = static_cast<int>(r / (gh * gw)), p = static_cast<int>(r - (long long)b * gh * gw);
        const int ph = p / gw, pw = p - ph * gw;
        const float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * C + c) * H + ph * P + ii) * W + pw * P + jj);
        *reinterpret_cast<uint2*>(out + r * K + k) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

// prefix tokens: out[b, t, :] = tok[t, :] + pos[t, :]  for t < T   (cls / dist tokens; models/deit.py:38-42)
__global__ void __launch_bounds__(256)
prefix_tokens_kernel(const float* __restrict__ tok, const float* __restrict__ pos, float* __restrict__ out, int B, int T,
                     long long tokens_per_image, int D) {
    const long long total = (long long)B * T * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = static_cast<int>(i % D);
        const long long bt = i / D;
        const int t = static_cast<int>(bt % T);
        const long long b = bt / T;
        out[(b * tokens_per_image + t) * D + c] = tok[(long long)t * D + c] + pos[(long long)t * D + c];
    }
}

static inline int ew_grid(long long work_items) {
    long long blocks = (work_items + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static inline int ln_grid(long long rows) {
    const long long need = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)sm_count() * 8;
    return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------------------------------------------------------
// Fused softmax cross-entropy (mean) + gradient + argmax accuracy of one batch of logits: replaces CrossEntropyLoss
// (utils_network.py:429-433, created main.py:244), its autograd and classification_count_correct
// (utils_network.py:85-95) without a host synchronisation. One CTA, one warp per row (round-robin); the per-warp
// partial losses are combined in a fixed order, so the result is deterministic.
//   out[0] = mean_r (logsumexp(x_r) - x_r[label_r]);  out[1] = #(argmax_c x_r[c] == label_r) (first maximum, as
//   torch.argmax);  dlogits[r, c] = (softmax(x_r)[c] - [c == label_r]) / rows
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
cross_entropy_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int rows,
                     int C, float* __restrict__ out, float* __restrict__ dlogits, long long ldd) {
    __shared__ float s_loss[32];
    __shared__ int s_correct[32];
    __shared__ int s_valid[32];
    __shared__ int s_bad[32];
    constexpr long long kIgnore = -100;   // nn.CrossEntropyLoss default ignore_index: such rows leave the mean
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // pass 0: rows that count (torch divides by the number of non-ignored targets); any other out-of-range label is an
    // error (torch device-asserts): the loss becomes NaN so that it cannot pass unnoticed
    int nvalid = 0, nbad = 0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const long long l = labels[r];
        if (l == kIgnore) continue;
        if (l < 0 || l >= C) ++nbad;
        ++nvalid;
    }
    nvalid = warp_sum(nvalid);
    nbad = warp_sum(nbad);
    if (lane == 0) { s_valid[warp] = nvalid; s_bad[warp] = nbad; }
    __syncthreads();
    nvalid = 0; nbad = 0;
    for (int w = 0; w < nwarps; ++w) { nvalid += s_valid[w]; nbad += s_bad[w]; }
    const float inv_rows = 1.0f / (float)nvalid;     // 0 valid rows: inf * 0 = NaN, as torch
    float loss = 0.f;
    int correct = 0;
    for (int r = warp; r < rows; r += nwarps) {
        const float* x = logits + (long long)r * ld;
        const long long label64 = labels[r];
        const bool ignored = label64 == kIgnore;
        const int label = (int)label64;
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = x[c];
            if (v > mx) { mx = v; arg = c; }   // strict: keeps the first maximum of this lane's columns
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) sum += __expf(x[c] - mx);
        sum = warp_sum(sum);
        const float lse = mx + __logf(sum);
        const float inv = 1.0f / sum;
        if (dlogits != nullptr) {
            float* d = dlogits + (long long)r * ldd;
            for (int c = lane; c < C; c += 32)
                d[c] = ignored ? 0.f : (__expf(x[c] - mx) * inv - (c == label ? 1.0f : 0.0f)) * inv_rows;
        }
        if (lane == 0) {
            if (!ignored) loss += lse - ((label >= 0 && label < C) ? x[label] : 0.f);
            correct += (arg == label) ? 1 : 0;
        }
    }
    if (lane == 0) { s_loss[warp] = loss; s_correct[warp] = correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f;
        int tc = 0;
        for (int w = 0; w < nwarps; ++w) { tl += s_loss[w]; tc += s_correct[w]; }
        out[0] = nbad > 0 ? __int_as_float(0x7fc00000) : tl * inv_rows;
        out[1] = (float)tc;
    }
}

}  // namespace vitk

using namespace vitk;

static int ln_fwd_impl(const float* x, long long x_stride, const float* weight, const float* bias, void* y_bf16,
                       float* y_f32, float* mean, float* rstd, long long rows, int D, float eps, void* stream) {
    if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 1024 || (x_stride % 4) != 0) return VITK_ERR_ARG;
    if (!x || !weight || !bias || (!y_bf16 && !y_f32) || !mean || !rstd) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto y = reinterpret_cast<__nv_bfloat16*>(y_bf16);
    const int chunks = (D + 127) / 128;
    const int grid = ln_grid(rows);
#define VITK_LN_FWD(C) \
    ln_fwd_kernel<C><<<grid, LN_WARPS * 32, 0, st>>>(x, x_stride, weight, bias, y, y_f32, mean, rstd, rows, D, eps)
    if (chunks <= 2) VITK_LN_FWD(2);
    else if (chunks <= 3) VITK_LN_FWD(3);
    else if (chunks <= 6) VITK_LN_FWD(6);
    else VITK_LN_FWD(8);
#undef VITK_LN_FWD
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y_bf16, float* mean,
                                  float* rstd, long long rows, int D, float eps, void* stream) {
    return ln_fwd_impl(x, D, weight, bias, y_bf16, nullptr, mean, rstd, rows, D, eps, stream);
}

extern "C" int vitk_layernorm_fwd_ex(const float* x, long long x_stride, const float* weight, const float* bias,
                                     void* y_bf16, float* y_f32, float* mean, float* rstd, long long rows, int D,
                                     float eps, void* stream) {
    return ln_fwd_impl(x, x_stride, weight, bias, y_bf16, y_f32, mean, rstd, rows, D, eps, stream);
}

static int ln_bwd_impl(const void* dy, int dy_is_f32, const float* x, long long x_stride, const float* weight,
                       const float* mean, const float* rstd, const float* dres, float* dx, long long dx_stride,
                       void* dx_bf16, const float* colscale, float* dweight, float* dbias, float* dxsum, long long rows,
                       int D, void* stream) {
    if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 1024 || (x_stride % 4) != 0 || (dx_stride % 4) != 0)
        return VITK_ERR_ARG;
    if (!dy || !x || !weight || !mean || !rstd) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
    const int chunks = (D + 127) / 128;
    // large problems: TMA-staged ring kernel (bulk copies need 16-byte row segments: D % 8 == 0 covers bf16 rows)
    static int ring_off = -1;
    if (ring_off < 0) { const char* e = getenv("VITK_LN_RING"); ring_off = (e != nullptr && e[0] == '0') ? 1 : 0; }
    if (!ring_off && rows >= 4096 && (D % 8) == 0 && D >= 128 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0) {
        const int cpad = chunks <= 2 ? 2 : chunks <= 3 ? 3 : chunks <= 6 ? 6 : 8;
        const int Dp = cpad * 128;
        // rows per stage: two when the rows are contiguous (three bulk copies per stage instead of per row). Measured,
        // rows x D = 25088 x 384: 52.3 us (1 row), 33.6 (2), 37.3 (4); 25216 x 768: 66.1 (1), 58.4 (2)
        const unsigned row_bytes = (unsigned)(D * 4 + D * 4 + D * (dy_is_f32 ? 4 : 2));
        int R = 1;
        {
            static int r_forced = -1;
            if (r_forced < 0) { const char* e = getenv("VITK_LN_RING_ROWS"); r_forced = e ? atoi(e) : 0; }
            if (x_stride == D && (dres == nullptr || dx_stride == D)) {
                R = r_forced > 0 ? r_forced : 2;
                if (R < 1) R = 1;
                if (R > 8) R = 8;
                while (R > 1 && (long long)R * row_bytes * 4 > 200 * 1024) --R;     // keep at least four stages
            }
        }
        const unsigned stage_bytes = ((unsigned)R * row_bytes + 127u) & ~127u;
        const long long fixed = (long long)Dp * 4 + 2 * 64 * 8 + 256;
        int stages = (int)((220 * 1024 - fixed) / stage_bytes);
        if (stages > 32) stages = 32;
        const long long part_bytes = (long long)LNR_CONSUMERS * 3 * Dp * 4;
        while ((long long)stages * stage_bytes < part_bytes) ++stages;  // the ring doubles as the reduction scratch
        const int smem = (int)((long long)stages * stage_bytes + Dp * 4 + 2 * stages * 8 + 128);
        if (smem <= 227 * 1024 && stages >= 4) {
            const int grid = (int)(rows < sm_count() ? rows : sm_count());
#define VITK_LN_RING_ONE(C, F)                                                                                        \
    do {                                                                                                              \
        static int attr_smem = 0;                                                                                     \
        if (attr_smem < smem) {                                                                                       \
            if (cudaFuncSetAttribute(ln_bwd_ring_kernel<C, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=  \
                cudaSuccess)                                                                                          \
                return VITK_ERR_CUDA;                                                                                 \
            attr_smem = smem;                                                                                         \
        }                                                                                                             \
        ln_bwd_ring_kernel<C, F><<<grid, LNR_THREADS, smem, st>>>(dy, x, x_stride, weight, mean, rstd, dres, dx,      \
                                                                  dx_stride, dxb, colscale, dweight, dbias, dxsum,   \
                                                                  rows, D, stages, R);                                \
    } while (0)
#define VITK_LN_RING(C)                           \
    do {                                          \
        if (dy_is_f32) VITK_LN_RING_ONE(C, true); \
        else VITK_LN_RING_ONE(C, false);          \
    } while (0)
            if (cpad == 2) VITK_LN_RING(2);
            else if (cpad == 3) VITK_LN_RING(3);
            else if (cpad == 6) VITK_LN_RING(6);
            else VITK_LN_RING(8);
#undef VITK_LN_RING
#undef VITK_LN_RING_ONE
            return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
        }
    }
    long long need = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)sm_count() * 2;
    const int grid = (int)(need < cap ? need : cap);
#define VITK_LN_BWD_ONE(C, F)                                                                                        \
    do {                                                                                                             \
        constexpr int smem = (1 + 3 * LN_WARPS) * (C) * 128 * 4;                                                     \
        static bool attr = false;                                                                                    \
        if (!attr) {                                                                                                 \
            if (cudaFuncSetAttribute(ln_bwd_kernel<C, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=      \
                cudaSuccess)                                                                                         \
                return VITK_ERR_CUDA;                                                                                \
            attr = true;                                                                                             \
        }                                                                                                            \
        ln_bwd_kernel<C, F><<<grid, LN_WARPS * 32, smem, st>>>(dy, x, x_stride, weight, mean, rstd, dres, dx,        \
                                                               dx_stride, dxb, colscale, dweight, dbias, dxsum, rows, D); \
    } while (0)
#define VITK_LN_BWD(C)                      \
    do {                                    \
        if (dy_is_f32) VITK_LN_BWD_ONE(C, true); \
        else VITK_LN_BWD_ONE(C, false);     \
    } while (0)
    if (chunks <= 2) VITK_LN_BWD(2);
    else if (chunks <= 3) VITK_LN_BWD(3);
    else if (chunks <= 6) VITK_LN_BWD(6);
    else VITK_LN_BWD(8);
#undef VITK_LN_BWD
#undef VITK_LN_BWD_ONE
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_layernorm_bwd(const void* dy_bf16, const float* x, const float* weight, const float* mean,
                                  const float* rstd, const float* dres, float* dx, void* dx_bf16,
                                  const float* colscale, float* dweight, float* dbias, long long rows, int D,
                                  void* stream) {
    return ln_bwd_impl(dy_bf16, 0, x, D, weight, mean, rstd, dres, dx, D, dx_bf16, colscale, dweight, dbias, nullptr,
                       rows, D, stream);
}

extern "C" int vitk_layernorm_bwd_ex(const void* dy, int dy_is_f32, const float* x, long long x_stride,
                                     const float* weight, const float* mean, const float* rstd, const float* dres,
                                     float* dx, long long dx_stride, void* dx_bf16, const float* colscale,
                                     float* dweight, float* dbias, float* dxsum, long long rows, int D, void* stream) {
    return ln_bwd_impl(dy, dy_is_f32, x, x_stride, weight, mean, rstd, dres, dx, dx_stride, dx_bf16, colscale, dweight,
                       dbias, dxsum, rows, D, stream);
}

extern "C" int vitk_colsum_bf16(const void* x_bf16, long long ldx, long long rows, int N, float* out, void* stream) {
    if (rows <= 0 || N <= 0 || (N % 8) != 0 || (ldx % 8) != 0 || !x_bf16 || !out) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int gx = (N + 255) / 256;
    long long gy = (4LL * sm_count() + gx - 1) / gx;
    const long long max_gy = (rows + 7) / 8;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    colsum_bf16_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16), ldx, rows,
                                                              N, out);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

namespace vitk {
__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
    const long long n8 = n / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = ld_nc_v4(x + i * 8);
        *reinterpret_cast<float4*>(y + i * 8) = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
        *reinterpret_cast<float4*>(y + i * 8 + 4) = make_float4(bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w));
    }
    if (blockIdx.x == 0)
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) y[i] = __bfloat162float(x[i]);
}
}  // namespace vitk

extern "C" int vitk_cast_bf16_f32(const void* x_bf16, float* y, long long n, void* stream) {
    if (n <= 0 || !x_bf16 || !y || ((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(y)) & 15))
        return VITK_ERR_ARG;
    vitk::cast_bf16_f32_kernel<<<ew_grid(n / 8 + 1), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_bf16), y, n);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_cast_f32_bf16(const float* x, void* y_bf16, long long n, void* stream) {
    if (n <= 0 || !x || !y_bf16) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (n / 8 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cast_f32_bf16_kernel<<<(int)blocks, 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y_bf16), n);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_cross_entropy(const float* logits, long long ld, const long long* labels, int rows, int C,
                                  float* out2, float* dlogits, long long ldd, void* stream) {
    if (rows <= 0 || C <= 0 || ld < C || !logits || !labels || !out2 || (dlogits && ldd < C)) return VITK_ERR_ARG;
    int threads = rows >= 32 ? 1024 : rows * 32;
    cross_entropy_kernel<<<1, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, ld, labels, rows, C, out2,
                                                                                     dlogits, ldd);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_colsum_f32(const float* x, long long ldx, long long rows, long long cols, float* out, void* stream) {
    if (rows <= 0 || cols <= 0 || (cols % 4) != 0 || (ldx % 4) != 0 || !x || !out) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long gx = (cols / 4 + 255) / 256;
    long long gy = (2LL * sm_count() + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy < 1) gy = 1;
    colsum_f32_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(x, ldx, rows, cols, out);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_colsum_prod(const float* a, long long lda, const void* b_bf16, long long ldb, long long rows, int N,
                                float* out, void* stream) {
    return vitk_colsum_prod_ex(a, lda, b_bf16, ldb, rows, N, out, nullptr, 1, stream);
}

extern "C" int vitk_colsum_prod_ex(const float* a, long long lda, const void* b_bf16, long long ldb, long long rows,
                                   int N, float* out, const float* rowscale, long long rows_per_sample, void* stream) {
    if (rows <= 0 || N <= 0 || (N % 8) != 0 || (lda % 4) != 0 || (ldb % 8) != 0 || !a || !b_bf16 || !out ||
        (rowscale != nullptr && rows_per_sample <= 0))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int gx = (N + 255) / 256;
    long long gy = (4LL * sm_count() + gx - 1) / gx;
    const long long max_gy = (rows + 7) / 8;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    colsum_prod_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(a, lda, reinterpret_cast<const __nv_bfloat16*>(b_bf16), ldb,
                                                              rows, N, out, rowscale, rows_per_sample);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_layerscale_bwd(const float* dy, long long lddy, const void* f_bf16, long long ldf, const float* gamma,
                                   const float* rowscale, long long rows_per_sample, long long rows, int N,
                                   void* out_bf16, float* dgamma, float* dbias, void* stream) {
    if (rows <= 0 || N <= 0 || (N % 8) != 0 || (lddy % 4) != 0 || (ldf % 8) != 0 || !dy || !f_bf16 || !gamma || !out_bf16 ||
        (rowscale != nullptr && rows_per_sample <= 0))
        return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int gx = (N + 255) / 256;
    long long gy = (4LL * sm_count() + gx - 1) / gx;
    const long long max_gy = (rows + 7) / 8;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    layerscale_bwd_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(dy, lddy, reinterpret_cast<const __nv_bfloat16*>(f_bf16),
                                                                 ldf, gamma, rowscale, rows_per_sample, rows, N,
                                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16), dgamma, dbias);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_scale_cast(const float* x, long long rows_per_group, long long group_stride, long long rows, int D,
                               const float* colscale, const float* rowscale, long long rows_per_sample, void* out_bf16,
                               void* stream) {
    if (rows <= 0 || D <= 0 || (D % 8) != 0 || rows_per_group <= 0 || (group_stride % 4) != 0 || !x || !out_bf16)
        return VITK_ERR_ARG;
    if (rowscale != nullptr && rows_per_sample <= 0) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    scale_cast_kernel<<<ew_grid(rows * (D / 8)), 256, 0, st>>>(x, rows_per_group, group_stride, rows, D, colscale, rowscale,
                                                              rows_per_sample > 0 ? rows_per_sample : 1,
                                                              reinterpret_cast<__nv_bfloat16*>(out_bf16));
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_patchify(const float* x, void* out_bf16, int B, int C, int H, int W, int P, void* stream) {
    if (B <= 0 || C <= 0 || P <= 0 || (P % 4) != 0 || H % P != 0 || W % P != 0 || !x || !out_bf16) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)B * (H / P) * (W / P) * (C * P * P / 4);
    patchify_kernel<<<ew_grid(total), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, C, H, W, P);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_prefix_tokens(const float* tok, const float* pos, float* out, int B, int T, long long tokens_per_image,
                                  int D, void* stream) {
    if (B <= 0 || T <= 0 || D <= 0 || !tok || !pos || !out) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    prefix_tokens_kernel<<<ew_grid((long long)B * T * D), 256, 0, st>>>(tok, pos, out, B, T, tokens_per_image, D);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_sgd_chunk_elems(void) { return SGD_CHUNK; }

extern "C" int vitk_sgd_momentum_multi(const void* table, const void* chunk_map, int num_chunks, float lr,
                                       float momentum, float grad_scale, int first_step, void* stream) {
    if (num_chunks < 0 || (num_chunks > 0 && (!table || !chunk_map))) return VITK_ERR_ARG;
    if (num_chunks == 0) return VITK_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    sgd_momentum_multi_kernel<<<num_chunks, 256, 0, st>>>(reinterpret_cast<const SgdTensor*>(table),
                                                          reinterpret_cast<const int2*>(chunk_map), lr, momentum,
                                                          grad_scale, first_step, nullptr);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_sgd_momentum_multi_hp(const void* table, const void* chunk_map, int num_chunks, const float* hyper,
                                          void* stream) {
    if (num_chunks < 0 || !hyper || (num_chunks > 0 && (!table || !chunk_map))) return VITK_ERR_ARG;
    if (num_chunks == 0) return VITK_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    sgd_momentum_multi_kernel<<<num_chunks, 256, 0, st>>>(reinterpret_cast<const SgdTensor*>(table),
                                                          reinterpret_cast<const int2*>(chunk_map), 0.f, 0.f, 1.f, 0,
                                                          hyper);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_adam_multi(const void* table, const void* chunk_map, int num_chunks, float* hyper, void* stream) {
    if (num_chunks < 0 || !hyper || (num_chunks > 0 && (!table || !chunk_map))) return VITK_ERR_ARG;
    if (num_chunks == 0) return VITK_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    adam_tick_kernel<<<1, 1, 0, st>>>(hyper);
    adam_multi_kernel<<<num_chunks, 256, 0, st>>>(reinterpret_cast<const AdamTensor*>(table),
                                                  reinterpret_cast<const int2*>(chunk_map), hyper);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_abi_version(void) { return 1; }
