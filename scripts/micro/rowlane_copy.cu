// Microbenchmark: fp32 y = x + 1 over [25216 x 768] with the GEMM epilogue's access pattern (lane == row, 64 B per
// lane per step, row stride 3072 B) against a row-contiguous pattern (a warp step covers 2 KB of one row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowlane_copy rowlane_copy.cu && ./rowlane_copy
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
constexpr int M = 25216, N = 768;
// MODE 0: lane == row, DEPTH chunks of 64 B in flight per lane; one warp owns 32 rows x SEG columns
template <int DEPTH, int SEG>
__global__ void __launch_bounds__(256) rowlane(const float* __restrict__ x, float* __restrict__ y, int nunits) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int CH = SEG / 16;  // 16-column chunks per unit
    for (int u = warp; u < nunits; u += nwarps) {
        const int rb = u / (N / SEG), cb = u % (N / SEG);
        const float* xr = x + (size_t)(rb * 32 + lane) * N + cb * SEG;
        float* yr = y + (size_t)(rb * 32 + lane) * N + cb * SEG;
        uint32_t q[DEPTH][16];
#pragma unroll
        for (int j = 0; j < DEPTH; ++j) { ldg256(xr + j * 16, q[j]); ldg256(xr + j * 16 + 8, q[j] + 8); }
#pragma unroll 1
        for (int c0 = 0; c0 < CH; c0 += DEPTH) {
#pragma unroll
            for (int j = 0; j < DEPTH; ++j) {
                const int c = c0 + j;
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(q[j][i]) + 1.0f);
                stg256(yr + c * 16, o); stg256(yr + c * 16 + 8, o + 8);
                if (c + DEPTH < CH) { ldg256(xr + (c + DEPTH) * 16, q[j]); ldg256(xr + (c + DEPTH) * 16 + 8, q[j] + 8); }
            }
        }
    }
}
// MODE 1: row-contiguous: a warp step covers 32 lanes x 32 B = 1 KB of one row
__global__ void __launch_bounds__(256) contig(const float* __restrict__ x, float* __restrict__ y, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t q[8];
        ldg256(x + i * 8, q);
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __float_as_uint(__uint_as_float(q[k]) + 1.0f);
        stg256(y + i * 8, q);
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f(i);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) f(i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20 * 1e3f;
}
int main() {
    float *x[4], *y[4];
    for (int i = 0; i < 4; ++i) { cudaMalloc(&x[i], (size_t)M * N * 4); cudaMalloc(&y[i], (size_t)M * N * 4); cudaMemset(x[i], 0, (size_t)M * N * 4); }
    const double bytes = 2.0 * M * N * 4;
    auto rep = [&](const char* name, float us) { printf("%-44s %7.1f us  %5.2f TB/s\n", name, us, bytes / us / 1e6); };
    rep("contiguous, 148x8 blocks", timeit([&](int i) { contig<<<148 * 8, 256>>>(x[i & 3], y[i & 3], (size_t)M * N / 8); }));
    rep("lane==row 64B/step, depth4, seg128, 148 blk x8w", timeit([&](int i) { rowlane<4, 128><<<148, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 128)); }));
    rep("lane==row 64B/step, depth4, seg128, 296 blk", timeit([&](int i) { rowlane<4, 128><<<296, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 128)); }));
    rep("lane==row 64B/step, depth4, seg128, 1184 blk", timeit([&](int i) { rowlane<4, 128><<<1184, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 128)); }));
    rep("lane==row, depth8 (512B/lane), seg128, 148 blk", timeit([&](int i) { rowlane<8, 128><<<148, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 128)); }));
    rep("lane==row, depth8, seg256, 148 blk", timeit([&](int i) { rowlane<8, 256><<<148, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 256)); }));
    rep("lane==row, depth2, seg128, 148 blk", timeit([&](int i) { rowlane<2, 128><<<148, 256>>>(x[i & 3], y[i & 3], (M / 32) * (N / 128)); }));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
