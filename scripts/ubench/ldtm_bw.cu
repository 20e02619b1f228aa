// Microbenchmark: TMEM read (tcgen05.ld) throughput per SM for different shapes / warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ uint32_t ld_tmem(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ld_tmem<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i];
    return x;
}
template <> __device__ __forceinline__ uint32_t ld_tmem<16>(uint32_t taddr) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= v[i];
    return x;
}
// two loads in flight before the wait
__device__ __forceinline__ uint32_t ld_tmem_2x32(uint32_t taddr) {
    uint32_t v[32], w[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
          "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
          "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
        : "r"(taddr + 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i] ^ w[i];
    return x;
}

// MODE 0: x32 + wait per load; 1: x16 + wait; 2: two x32 then wait
template <int MODE>
__global__ void ldtm_kernel(int iters, uint32_t* out, long long* cycles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)(((it * 64) + (warp >> 2) * 64) & 511) & ~63u;
        if (MODE == 0) { acc ^= ld_tmem<32>(lane_base + col); acc ^= ld_tmem<32>(lane_base + col + 32); }
        if (MODE == 1) { acc ^= ld_tmem<16>(lane_base + col); acc ^= ld_tmem<16>(lane_base + col + 16); acc ^= ld_tmem<16>(lane_base + col + 32); acc ^= ld_tmem<16>(lane_base + col + 48); }
        if (MODE == 2) { acc ^= ld_tmem_2x32(lane_base + col); }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
    }
}

template <int MODE> void run(int warps, int grid) {
    const int iters = 2000;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)grid * warps * 32 * 4);
    cudaMalloc(&cyc, grid * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    ldtm_kernel<MODE><<<grid, warps * 32>>>(iters, out, cyc);
    cudaEventRecord(e0);
    ldtm_kernel<MODE><<<grid, warps * 32>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // bytes per iteration per warp: 64 columns x 32 lanes x 4 B = 8 KB
    const double bytes = (double)iters * warps * 8192.0;
    printf("mode %d warps %2d grid %3d: %s  %.1f us, %lld cycles, %.1f B/cycle/SM, %.1f cycles per 4KB LDTM-equivalent per warp\n", MODE, warps, grid,
           cudaGetErrorString(e), ms * 1e3, c, bytes / (double)c, (double)c / (iters * 2.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int grid : {1, 148}) {
        for (int warps : {4, 8, 16}) { run<0>(warps, grid); run<1>(warps, grid); run<2>(warps, grid); }
    }
    return 0;
}
