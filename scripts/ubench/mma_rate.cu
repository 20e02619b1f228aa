// Microbenchmark: tcgen05.mma issue / execution rate on one SM — SS vs TS (A from TMEM), N = 64/128/256, one or two
// issuing threads, same or alternating accumulators.  cycles per MMA = (commit arrival - first issue) / count.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {   // A, B fp16, D fp32, K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// TS: A from TMEM; N: MMA N; ALT: accumulators used round-robin (1 or 2); issuers: 1..4 threads (lane 0 of warps 0..3), each with
// its own accumulator columns.  The issue loop is fully unrolled with precomputed operands (a single thread retires a
// dependent ALU instruction every ~4 cycles, so address arithmetic in the loop would dominate).
template <int TS, int N, int ALT>
__global__ void mma_kernel(int issuers, int count, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    if (warp < issuers && lane == 0) {
        constexpr uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t adesc = make_sw128_desc(smem_u32(smem));                  // 128 rows x 128 B
        const uint64_t bdesc = make_sw128_desc(smem_u32(smem + 32768));          // up to 256 rows x 128 B
        const uint32_t a_tmem = base + 448 + warp * 16;
        const uint32_t d0 = base + (uint32_t)((warp * ALT * N) % 448), d1 = ALT == 2 ? d0 + N : d0;
        // first MMAs overwrite
        if (TS) { mma_ts(d0, a_tmem, bdesc, idesc, 0u); mma_ts(d1, a_tmem, bdesc, idesc, 0u); }
        else { mma_ss(d0, adesc, bdesc, idesc, 0u); mma_ss(d1, adesc, bdesc, idesc, 0u); }
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 4) {
            if (TS) {
                mma_ts(d0, a_tmem + 0, bdesc + 0, idesc, 1u); mma_ts(d1, a_tmem + 8, bdesc + 2, idesc, 1u);
                mma_ts(d0, a_tmem + 0, bdesc + 4, idesc, 1u); mma_ts(d1, a_tmem + 8, bdesc + 6, idesc, 1u);
            } else {
                mma_ss(d0, adesc + 0, bdesc + 0, idesc, 1u); mma_ss(d1, adesc + 2, bdesc + 2, idesc, 1u);
                mma_ss(d0, adesc + 4, bdesc + 4, idesc, 1u); mma_ss(d1, adesc + 6, bdesc + 6, idesc, 1u);
            }
        }
        tc_commit(&bars[warp]);
        const long long t1 = clock64();
        while (!mbar_try_wait(&bars[warp], 0)) {}
        const long long t2 = clock64();
        out[warp * 2 + 0] = t1 - t0;
        out[warp * 2 + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
    }
}

template <int TS, int N, int ALT> void run(int issuers, long long* out) {
    const int count = 1024;
    cudaFuncSetAttribute(mma_kernel<TS, N, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaMemset(out, 0, 64);
    mma_kernel<TS, N, ALT><<<1, 128, 100 * 1024>>>(issuers, count, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("%s N=%3d accumulators=%d issuers=%d: %s  per issuer: issue %.1f cyc/MMA, complete %.1f cyc/MMA; all issuers together %.1f cyc/MMA (nominal %d)\n",
           TS ? "TS" : "SS", N, ALT, issuers, cudaGetErrorString(e), (double)h[0] / count, (double)h[1] / count,
           (double)h[2 * (issuers - 1) + 1] / count / issuers, N / 2);
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    for (int issuers : {1, 2, 4}) {
        run<0, 64, 1>(issuers, out); run<0, 64, 2>(issuers, out); run<0, 128, 1>(issuers, out);
        run<1, 64, 1>(issuers, out); run<1, 64, 2>(issuers, out); run<1, 128, 1>(issuers, out);
        if (issuers <= 2) { run<0, 128, 2>(issuers, out); run<1, 128, 2>(issuers, out); }
        if (issuers == 1) { run<0, 256, 1>(issuers, out); run<1, 256, 1>(issuers, out); }
    }
    return 0;
}
