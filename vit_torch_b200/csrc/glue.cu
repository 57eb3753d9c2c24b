// Memory-bound glue kernels of the ViT block: LayerNorm fwd/bwd (one warp per row, 128-bit loads, shuffle
// reductions), bias-gradient column sums, dtype casts. All HBM-roofline kernels: no smem staging needed because
// every element is touched once; grids are multiples of the SM count with grid-stride loops.
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/vitk.h"

namespace vitk {

constexpr int LN_WARPS = 8;

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: x fp32 [rows, D] -> y bf16, mean/rstd fp32.    (nn.LayerNorm, eps inside sqrt, biased variance)
// ------------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              long long rows, int D, float eps) {
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
        const float* xr = x + row * D;
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
        __nv_bfloat16* yr = y + row * D;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float o0 = (v[i].x - mean) * rstd * wv[i].x + bv[i].x;
                const float o1 = (v[i].y - mean) * rstd * wv[i].y + bv[i].y;
                const float o2 = (v[i].z - mean) * rstd * wv[i].z + bv[i].z;
                const float o3 = (v[i].w - mean) * rstd * wv[i].w + bv[i].w;
                *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward fused with residual-gradient add and the bf16 copy the next GEMM consumes.
//   dx = dres + rstd * (dy*w - mean_D(dy*w) - xhat * mean_D(dy*w*xhat));  dw += sum_rows dy*xhat;  db += sum_rows dy
// ------------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dres,
              float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16, const float* __restrict__ colscale,
              float* __restrict__ dweight, float* __restrict__ dbias, long long rows, int D) {
    __shared__ float red[LN_WARPS * MAXC * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 wv[MAXC], dwa[MAXC], dba[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = (i * 32 + lane) * 4;
        wv[i] = (c < D) ? __ldg(reinterpret_cast<const float4*>(w + c)) : make_float4(0, 0, 0, 0);
        dwa[i] = make_float4(0, 0, 0, 0);
        dba[i] = make_float4(0, 0, 0, 0);
    }
    const float inv_d = 1.0f / static_cast<float>(D);
    for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float* xr = x + row * D;
        const __nv_bfloat16* dyr = dy + row * D;
        float4 xh[MAXC], g[MAXC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float4 xv = *reinterpret_cast<const float4*>(xr + c);
                const uint2 dv = *reinterpret_cast<const uint2*>(dyr + c);
                const float d0 = bf16_lo(dv.x), d1 = bf16_hi(dv.x), d2 = bf16_lo(dv.y), d3 = bf16_hi(dv.y);
                xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                                    (xv.w - mean) * rstd);
                g[i] = make_float4(d0 * wv[i].x, d1 * wv[i].y, d2 * wv[i].z, d3 * wv[i].w);
                dwa[i].x += d0 * xh[i].x; dwa[i].y += d1 * xh[i].y; dwa[i].z += d2 * xh[i].z; dwa[i].w += d3 * xh[i].w;
                dba[i].x += d0; dba[i].y += d1; dba[i].z += d2; dba[i].w += d3;
                s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            } else {
                xh[i] = make_float4(0, 0, 0, 0);
                g[i] = make_float4(0, 0, 0, 0);
            }
        }
        const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float4 o = make_float4(rstd * (g[i].x - c1 - xh[i].x * c2), rstd * (g[i].y - c1 - xh[i].y * c2),
                                       rstd * (g[i].z - c1 - xh[i].z * c2), rstd * (g[i].w - c1 - xh[i].w * c2));
                if (dres != nullptr) {
                    const float4 r = *reinterpret_cast<const float4*>(dres + row * D + c);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * D + c) = o;
                if (dx_bf16 != nullptr) {
                    if (colscale != nullptr) {
                        const float4 cs = __ldg(reinterpret_cast<const float4*>(colscale + c));
                        o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
                    }
                    *reinterpret_cast<uint2*>(dx_bf16 + row * D + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                }
            }
        }
    }
    // block reduction of the per-warp dweight / dbias partials, then one atomic per column per block
    const int Dp = MAXC * 128;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dweight : dbias;
        if (dst == nullptr) continue;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (i * 32 + lane) * 4;
            *reinterpret_cast<float4*>(&red[warp * Dp + c]) = pass == 0 ? dwa[i] : dba[i];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < LN_WARPS; ++k) s += red[k * Dp + c];
            atomicAdd(dst + c, s);
        }
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

static inline int ln_grid(long long rows) {
    const long long need = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)sm_count() * 8;
    return (int)(need < cap ? need : cap);
}

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y_bf16, float* mean,
                                  float* rstd, long long rows, int D, float eps, void* stream) {
    if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 1024) return VITK_ERR_ARG;
    if (!x || !weight || !bias || !y_bf16 || !mean || !rstd) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = ln_grid(rows);
    auto y = reinterpret_cast<__nv_bfloat16*>(y_bf16);
    const int chunks = (D + 127) / 128;
    if (chunks <= 2) ln_fwd_kernel<2><<<grid, LN_WARPS * 32, 0, st>>>(x, weight, bias, y, mean, rstd, rows, D, eps);
    else if (chunks <= 3) ln_fwd_kernel<3><<<grid, LN_WARPS * 32, 0, st>>>(x, weight, bias, y, mean, rstd, rows, D, eps);
    else if (chunks <= 6) ln_fwd_kernel<6><<<grid, LN_WARPS * 32, 0, st>>>(x, weight, bias, y, mean, rstd, rows, D, eps);
    else ln_fwd_kernel<8><<<grid, LN_WARPS * 32, 0, st>>>(x, weight, bias, y, mean, rstd, rows, D, eps);
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
}

extern "C" int vitk_layernorm_bwd(const void* dy_bf16, const float* x, const float* weight, const float* mean,
                                  const float* rstd, const float* dres, float* dx, void* dx_bf16,
                                  const float* colscale, float* dweight, float* dbias, long long rows, int D,
                                  void* stream) {
    if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 1024) return VITK_ERR_ARG;
    if (!dy_bf16 || !x || !weight || !mean || !rstd) return VITK_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long need = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)sm_count() * 4;
    const int grid = (int)(need < cap ? need : cap);
    auto dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
    auto dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
    const int chunks = (D + 127) / 128;
#define VITK_LN_BWD(C)                                                                                             \
    ln_bwd_kernel<C><<<grid, LN_WARPS * 32, 0, st>>>(dy, x, weight, mean, rstd, dres, dx, dxb, colscale, dweight, \
                                                     dbias, rows, D)
    if (chunks <= 2) VITK_LN_BWD(2);
    else if (chunks <= 3) VITK_LN_BWD(3);
    else if (chunks <= 6) VITK_LN_BWD(6);
    else VITK_LN_BWD(8);
#undef VITK_LN_BWD
    return cudaGetLastError() == cudaSuccess ? VITK_OK : VITK_ERR_CUDA;
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

extern "C" int vitk_abi_version(void) { return 1; }
