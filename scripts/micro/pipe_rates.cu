// Micro-benchmark of the three per-SM rates the attention / talking-heads kernels are bounded by (B200, sm_100a):
//   (1) tcgen05.ld throughput (TMEM -> registers) with 1, 2, 4, 8 warps per SM,
//   (2) MUFU ex2 throughput with 4, 8, 16 warps per SM,
//   (3) mma.sync.m16n8k8 tf32 throughput with 4, 8, 16 warps per SM (independent and dependent accumulators).
// One CTA per SM (grid = #SM), SM clock via clock64(). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// scripts/micro/bin/pipe_rates scripts/micro/pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) tmem_ld_rate(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp < nwarps) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t r[32];
            const uint32_t a = base + ((it & 7) * 32) + (warp >= 4 ? 256 : 0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
                "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(a));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i += 8) acc += __uint_as_float(r[i]);
        }
        t1 = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}

// same, but two loads in flight before each wait (x32 + x32)
__global__ void __launch_bounds__(256, 1) tmem_ld_rate2(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp < nwarps) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t r[64];
            const uint32_t a = base + ((it & 3) * 64) + (warp >= 4 ? 256 : 0);
#pragma unroll
            for (int h = 0; h < 2; ++h)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
                    "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[32 * h + 0]), "=r"(r[32 * h + 1]), "=r"(r[32 * h + 2]), "=r"(r[32 * h + 3]), "=r"(r[32 * h + 4]),
                      "=r"(r[32 * h + 5]), "=r"(r[32 * h + 6]), "=r"(r[32 * h + 7]), "=r"(r[32 * h + 8]), "=r"(r[32 * h + 9]),
                      "=r"(r[32 * h + 10]), "=r"(r[32 * h + 11]), "=r"(r[32 * h + 12]), "=r"(r[32 * h + 13]),
                      "=r"(r[32 * h + 14]), "=r"(r[32 * h + 15]), "=r"(r[32 * h + 16]), "=r"(r[32 * h + 17]),
                      "=r"(r[32 * h + 18]), "=r"(r[32 * h + 19]), "=r"(r[32 * h + 20]), "=r"(r[32 * h + 21]),
                      "=r"(r[32 * h + 22]), "=r"(r[32 * h + 23]), "=r"(r[32 * h + 24]), "=r"(r[32 * h + 25]),
                      "=r"(r[32 * h + 26]), "=r"(r[32 * h + 27]), "=r"(r[32 * h + 28]), "=r"(r[32 * h + 29]),
                      "=r"(r[32 * h + 30]), "=r"(r[32 * h + 31])
                    : "r"(a + 32 * h));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 64; i += 8) acc += __uint_as_float(r[i]);
        }
        t1 = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}

__global__ void __launch_bounds__(512, 1) mufu_rate(int iters, long long* out, float* sink) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 123.456f) sink[0] = s;
}

// FFMA2 (packed fp32x2) rate: 8 independent chains per thread
__global__ void __launch_bounds__(512, 1) ffma2_rate(int iters, long long* out, float* sink) {
    unsigned long long x[8];
    const unsigned long long c = 0x3f8000003f800000ull;   // (1.0f, 1.0f)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = c;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[i]) : "l"(c));
    }
    const long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 12345ull) sink[0] = 1.f;
}

template <bool DEP>
__global__ void __launch_bounds__(512, 1) mma_tf32_rate(int iters, long long* out, float* sink) {
    float d[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* dd = DEP ? d[0] : d[j];
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[j][0] + d[j][3];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 123.456f) sink[0] = s;
}

int main() {
    long long* out;
    float* sink;
    cudaMalloc(&out, 8);
    cudaMalloc(&sink, 4);
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long h = 0;
    const int iters = 4096;
    for (int nw : {1, 2, 4, 8}) {
        tmem_ld_rate<<<nsm, 256>>>(nw, iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("tcgen05.ld 32x32b.x32 (wait each): %d warps/SM: %.1f clk per load per warp -> %.1f B/clk/SM\n", nw,
               (double)h / iters, 4096.0 * nw * iters / (double)h);
        tmem_ld_rate2<<<nsm, 256>>>(nw, iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("tcgen05.ld 2 x (32x32b.x32) per wait:  %d warps/SM: %.1f clk per pair per warp -> %.1f B/clk/SM\n", nw,
               (double)h / iters, 8192.0 * nw * iters / (double)h);
    }
    for (int nt : {128, 256, 512}) {
        mufu_rate<<<nsm, nt>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("ex2.approx: %d threads/SM: %.2f results/clk/SM\n", nt, 8.0 * iters * nt / (double)h);
        ffma2_rate<<<nsm, nt>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("fma.rn.f32x2: %d threads/SM: %.2f packed instr-lanes/clk/SM (x2 FMA each)\n", nt, 8.0 * iters * nt / (double)h);
        mma_tf32_rate<false><<<nsm, nt>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("mma.sync m16n8k8 tf32 (4 independent accumulators): %d warps/SM: %.2f clk per mma per SMSP\n", nt / 32,
               (double)h / (4.0 * iters * (nt / 32) / 4.0));
        mma_tf32_rate<true><<<nsm, nt>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("mma.sync m16n8k8 tf32 (one dependent chain):        %d warps/SM: %.2f clk per mma per warp\n", nt / 32,
               (double)h / (4.0 * iters));
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
