// Shared helpers for the sm_100a kernels behind include/oov_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/oov_b200.h"

namespace oov {

extern std::atomic<uint64_t> g_launches;
void set_error(const char* fmt, ...);
int num_sms();
int cur_device();        // current CUDA device clamped to [0, 64): index of per-device caches (function attributes and
                         // occupancy are per device / context, a process may drive several GPUs)

#define OOV_REQUIRE(cond, code, ...)         \
    do {                                     \
        if (!(cond)) {                       \
            ::oov::set_error(__VA_ARGS__);   \
            return (code);                   \
        }                                    \
    } while (0)

// Counts the launch and converts a launch-time error into OOV_ERR_CUDA.
#define OOV_LAUNCH_CHECK(name)                                                             \
    do {                                                                                   \
        ::oov::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            ::oov::set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(_e)); \
            return OOV_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
__device__ __forceinline__ bool aligned_dev(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------
// dtype-generic element access (tables / outputs are fp32 or bf16)
// ------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_elem(const void* base, int dtype, int64_t idx) {
    return dtype == OOV_F32 ? reinterpret_cast<const float*>(base)[idx]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}
__device__ __forceinline__ void store_elem(void* base, int dtype, int64_t idx, float v) {
    if (dtype == OOV_F32) reinterpret_cast<float*>(base)[idx] = v;
    else reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}
static inline size_t dtype_size(int dtype) { return dtype == OOV_F32 ? 4 : 2; }
static inline bool dtype_ok(int dtype) { return dtype == OOV_F32 || dtype == OOV_BF16; }

// Copy one D-element row between (possibly different) dtypes; `lane`/`nlanes` threads cooperate.
// Uses 16-byte accesses when both sides are 16-byte aligned and same dtype.
__device__ __forceinline__ void copy_row(const void* src, int sdt, void* dst, int ddt, int D, int lane, int nlanes) {
    if (sdt == ddt) {
        const size_t bytes = (size_t)D * (sdt == OOV_F32 ? 4 : 2);
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0) {
            const int4* s = reinterpret_cast<const int4*>(src);
            int4* d = reinterpret_cast<int4*>(dst);
            for (int i = lane; i < (int)(bytes >> 4); i += nlanes) d[i] = __ldg(s + i);
            return;
        }
    }
    if (sdt == OOV_F32 && ddt == OOV_BF16 && (D & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(dst) & 7)) == 0) {
        // fp32 table row -> bf16 table row (the in-vocab half of every assembled bf16 item table): 16 B in, 8 B out per lane
        const float4* s = reinterpret_cast<const float4*>(src);
        uint2* d = reinterpret_cast<uint2*>(dst);
        for (int i = lane; i < (D >> 2); i += nlanes) {
            const float4 v = __ldg(s + i);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            d[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
        return;
    }
    for (int i = lane; i < D; i += nlanes) store_elem(dst, ddt, i, load_elem(src, sdt, i));
}

// feature row of an id (training mode de-pads, lsh_embedder.py:153-155)
__device__ __forceinline__ int64_t feature_row(int64_t id, int64_t prime_pad) {
    return (prime_pad > 0 && id >= prime_pad) ? id - prime_pad : id;
}

// ------------------------------------------------------------------------------------
// Total order used by every top-k in this library: (score desc, id asc), NaN first
// (torch.topk ranks NaN above all numbers).  key64 = (ordered_float << 32) | ~id  (max = best)
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_order_key(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
    if (f != f) return 0xFFFFFFFFu;
#else
    memcpy(&u, &f, 4);
    if (f != f) return 0xFFFFFFFFu;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float float_from_order_key(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    if (k == 0xFFFFFFFFu) u = 0x7FC00000u;   // NaN
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
// Optional output mode of the final per-user merge: packed 8-byte candidates in GLOBAL item ids instead of (score, id)
// arrays — what a row shard laid out [ids lo0.. (n0 rows) | ids lo1..] sends to the other ranks (sharded.py).
struct KeyOut {
    unsigned long long* keys = nullptr;     // [Q, k]; nullptr: write out_scores / out_idx
    int64_t n0 = 0, lo0 = 0, lo1 = 0;       // local row r -> r < n0 ? lo0 + r : lo1 + (r - n0)
};
__device__ __forceinline__ unsigned long long make_key64(float score, uint32_t local_idx) {
    return ((unsigned long long)float_order_key(score) << 32) | (unsigned long long)(~local_idx);
}
__device__ __forceinline__ uint32_t key64_idx(unsigned long long k) { return ~(uint32_t)(k & 0xFFFFFFFFull); }
__device__ __forceinline__ float key64_score(unsigned long long k) { return float_from_order_key((uint32_t)(k >> 32)); }

}  // namespace oov
