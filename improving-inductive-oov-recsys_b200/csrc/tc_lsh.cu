// LSH OOV embedder on the tensor cores: sign-projection GEMM -> signs -> multi-hot x bucket-table GEMM -> mean.
//
// Replaces inductive/torch_hash.py:55-60 (R = X P^T, bit = !(R < 0)) and inductive/lsh_embedder.py:141-179
// (out = (H W) / H.sum(1)) in ONE kernel: neither R [n, B] fp32 nor H [n, B] ever exist in HBM or shared memory.
//
// Exact signs from fp16 tensor cores.  Only the SIGN of x . p matters, so both operands are rescaled freely: every
// plane to length 2^8 (p^ = 256 p / |p|), every feature row by a power of two (largest |x_i| -> [2^13, 2^14)), which
// keeps every fp16 piece in the normal range.  With x = x0 + x1 (+ <= 2^-22 |x|) and p^ = p0 + p1 (+ <= 2^-22 |p^|)
//     A' = [x0 | x1] in TMEM,   B' row = [p0 | p1],   R' = x0 p0 + x1 p0 + x0 p1      (six K = 16 steps, x0 reused)
// and  |R' - x . p^| <= 3 * 2^-22 |x||p^| + the tensor core's accumulation error.  A projection closer to zero than
// 2^-18 |x||p^| (about 2 in 100 000) is NOT trusted: its (row, plane) goes into a shared-memory queue and the worker warps go on.  At
// the end of the row tile all 512 worker threads re-evaluate the queued projections with the fp32 FMA chain of the
// CUDA-core path (csrc/lsh.cu: the original planes, f ascending) — so both paths give identical bits and count the same
// |R| < tie_eps events — and the few signs that really differ are applied to the finished accumulator as rank-one
// corrections (+-2 W[b, :]).  No warp ever leaves the pipeline for a long recomputation (the round-1 kernel lost 80 % of
// its time to exactly that: one slow lane held up all sixteen warps of every hand-over).
//
// Second GEMM without masks: the epilogue writes S' = 2H - 1 (+-1 in fp16: the sign bit of R under a constant), so
//     S' W = 2 H W - colsum(W)   ->   out = (acc + colsum(W)) / (sum(S') + B)        [sum(S') + B = 2 count]
// with 0 / 0 -> NaN like lsh_embedder.py:158.  sum(S') is accumulated by the workers with packed-half adds on the very
// words they store (exact: |sum| <= 2048).  W is split into fp16 hi (+ lo for fp32 outputs) pieces.
//
// Operands: B' (16 KB per 128 planes) is loaded ONCE per CTA when it fits its 8 stages (B <= 1024), the transposed
// bucket table (16 KB per 128 planes and piece) streams through a 4-stage TMA ring (resident when it fits); otherwise the
// same stages work as TMA rings.  (Streaming both per row tile would need ~50 B/clk/SM of L2 bandwidth at full speed;
// the chip delivers ~42.)
//
// Per CTA (640 threads), persistent over 128-row tiles of the id list (tiles without OOV ids are plain row copies and
// skip the GEMMs — one flag byte per tile, written by lsh_flags_kernel):
//   warps 0-15  workers : per 128-plane N tile (lane = row, warp = 32 of the 128 columns): tcgen05.ld the projections,
//                         min|R| tree against the near-zero threshold, two instructions per pair of scores (PRMT +
//                         LOP3) to form the fp16 +-1 words, one HADD2 per word for the count, tcgen05.st them back
//                         into TENSOR MEMORY as the A operand of the second GEMM.  Everything rare (a near-zero
//                         projection, planes >= B, the caller wants the bits) is one out-of-line call, so the loop stays
//                         a few hundred bytes of code (the inlined version lost a third of its time to it: instruction
//                         fetch stalls and the slowest of sixteen warps holding up every hand-over).
//   warp 16     TMA     : B' tiles and the transposed bucket-table tiles
//   warp 17     MMA 1   : GEMM1 (TS: A' from TMEM, M128 N128 K16 x 6) into one of two TMEM accumulators
//   warp 18     TMEM alloc, then fixer: re-evaluates the queued near-zero projections of a finished row tile while the
//                         workers go on
//   warp 19     MMA 2   : GEMM2 (TS: A = S' from TMEM, M128 N64 K16 x 8 per piece), accumulating over all N tiles.
// TMEM columns: 0-255 projections (2 buffers), 256-319 S' W, 320-447 S' (2 buffers), 448-479 A'.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {
namespace tc {

constexpr int L_BM = 128;                 // rows per tile
constexpr int L_BN = 128;                 // planes per N tile
constexpr int L_FCH = 32;                 // features per K chunk (K' = 3 * 32 = 96 fp16 per chunk)
constexpr int L_FMAX = 64;                // features: one or two chunks (A' of two chunks fills tensor-memory columns 448-511)
constexpr int L_DMAX = 64;
constexpr int L_BSTAGES = 8, L_WSTAGES = 4;   // stages of the B' / bucket-table rings (resident when everything fits)
constexpr int L_WORKERS = 16;
constexpr int W_TMA = 16, W_MMA1 = 17, W_FIXER = 18, W_MMA2 = 19;
constexpr int L_THREADS = 20 * 32;                             // 640 (a 21st warp would cost 16 registers per thread: 6 warps on one scheduler)
constexpr int L_BT_BYTES = L_BN * 128;                         // 16 KB: row r = [p0 | p1] of plane r of one N tile
constexpr int L_WT_BYTES = 2 * L_DMAX * 128;                   // 16 KB: 64 d-rows x 128 planes (two 64-plane K blocks)
constexpr int L_QCAP = 1024;                                   // queued near-zero projections per row tile
constexpr int L_TAIL_BYTES = 2 * 4 * L_BM * 4 + 4 * L_BM * 4 + 2 * L_BM * 8 + 4 * L_QCAP * 4 + L_DMAX * 4 + 1024;
constexpr int L_SMEM = 1024 + L_BSTAGES * L_BT_BYTES + L_WSTAGES * L_WT_BYTES + L_TAIL_BYTES;
static_assert(L_SMEM <= 232448, "tc_lsh shared memory");
constexpr float L_PSCALE = 256.f;                 // length of the rescaled planes
constexpr float L_NEAR_REL = 3.814697265625e-6f * L_PSCALE;  // 2^-18 |p^|: |R' - x.p^| stays below a quarter of this times |x|

struct LshParams {
    const float* feat; int64_t n_feat_rows; int F; int FC;   // FC = ceil(F / 32) feature chunks
    const float* planes; int B; int NT;               // NT = ceil(B / 128)
    const int64_t* ids; int64_t ids_stride; int64_t n; int64_t n_old; int64_t prime_pad;
    const void* iv_table; int iv_dtype;
    void* out; int out_dtype; int64_t out_stride; int D;
    int wsplit;                                        // 1: fp16 bucket table, 2: hi + lo
    int res_b, res_w;                                  // operand tiles resident in shared memory (loaded once per CTA)
    float tie_eps;
    uint32_t* bits_out; int words;
    unsigned long long* tie_count;
    const float* pn_min;                               // smallest non-zero plane norm (device scalar written by the pack kernel)
    const float* wsum;                                 // [64] column sums of the packed bucket table
    const __half* Wt; int64_t nb;                      // packed bucket table [wsplit * 64, nb] (rank-one sign corrections)
    const uint8_t* flags;                              // [ceil(n / 128)] 1 = the row tile holds an OOV id (lsh_flags_kernel)
    const float* cast_src; __nv_bfloat16* cast_dst; int64_t cast_n8;   // side job of the TMA warp: bf16(cast_src[0 .. 8 cast_n8))
    unsigned long long* trace;                         // profiling only (OOV_LSH_TRACE_PTR): [4 roles][4096] event << 56 | clock of CTA 0
};
// one timestamp of CTA 0 (roles: 0 MMA1, 1 MMA2, 2 worker warp 0, 3 worker warp 15); a no-op unless a trace buffer is set
#ifndef OOV_TRACE_W2
#define OOV_TRACE_W2 (L_WORKERS - 1)      /* second traced worker warp (profiling builds: -DOOV_TRACE_W2=n) */
#endif
#ifdef OOV_LSH_TRACE
#define LTRACE(role, ev)                                                                                         \
    do {                                                                                                         \
        if (p.trace != nullptr && blockIdx.x == 0 && trace_n < 4096)                                              \
            p.trace[(role) * 4096 + trace_n++] = ((unsigned long long)(ev) << 56) | ((unsigned long long)(trace_tag & 255) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); \
    } while (0)
#else
#define LTRACE(role, ev) do { (void)trace_n; } while (0)
#endif

// ---------------------------------------------------------------- operand packing (once per call)
// Bp [NT*FC*128, 64] fp16: row (nt * FC + c) * 128 + b % 128 = [p0 | p1] (32 columns each) of features 32 c .. 32 c + 31 of p^ = 256 p / |p|
//                      (zero for f >= F, b >= B and planes without a finite non-zero norm: those are always re-evaluated)
// Wt [wsplit*64, NT*128] fp16: rows s*64 + d = piece s of W[., d]; zero padding
__global__ void lsh_pack_kernel(const float* __restrict__ planes, int B, int F, int NT, const void* __restrict__ W, int w_dtype,
                                int D, int wsplit, __half* __restrict__ Bp, __half* __restrict__ Wt,
                                float* __restrict__ pn_min) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = (int64_t)NT * L_BN;
    const int FC = (F + L_FCH - 1) / L_FCH;
    if (t < nb * FC * 32) {
        // row rb of Bp = row (b % 128) of the 128-plane tile nt = b / 128, feature chunk c: tiles are laid out (nt, c)
        const int64_t rb = t >> 5;
        const int tile = (int)(rb / L_BN), bl = (int)(rb % L_BN);
        const int b = (tile / FC) * L_BN + bl, f = (tile % FC) * L_FCH + (int)(t & 31);
        float v = 0.f;
        if (b < B) {
            float s2 = 0.f;
            for (int j = 0; j < F; ++j) { const float pj = planes[(size_t)b * F + j]; s2 = fmaf(pj, pj, s2); }
            const float nrm = sqrtf(s2);
            if (nrm > 0.f && nrm < INFINITY) {
                if (f < F) v = planes[(size_t)b * F + f] / nrm * L_PSCALE;
                if (f == 0) atomicMin(reinterpret_cast<unsigned int*>(pn_min), __float_as_uint(nrm));
            }
        }
        const __half p0 = __float2half_rn(v);
        Bp[(size_t)rb * 64 + (t & 31)] = p0;
        Bp[(size_t)rb * 64 + 32 + (t & 31)] = __float2half_rn(v - __half2float(p0));
    }
    if (t < nb * L_DMAX) {
        const int d = (int)(t / nb);
        const int64_t b = t - (int64_t)d * nb;
        float w = 0.f;
        if (b < B && d < D) w = load_elem(W, w_dtype, b * D + d);
        const __half hi = __float2half_rn(w);
        Wt[(size_t)d * nb + b] = hi;
        if (wsplit == 2) Wt[(size_t)(L_DMAX + d) * nb + b] = __float2half_rn(w - __half2float(hi));
    }
}
// wsum[d] = sum over b of the packed pieces of W[b, d], fixed order (one warp per column, fp32 tree over 32 partial sums)
__global__ void lsh_wsum_kernel(const __half* __restrict__ Wt, int64_t nb, int wsplit, float* __restrict__ wsum) {
    const int d = blockIdx.x, lane = threadIdx.x;
    float s = 0.f;
    for (int64_t b = lane; b < nb; b += 32) {
        float w = __half2float(Wt[(size_t)d * nb + b]);
        if (wsplit == 2) w += __half2float(Wt[(size_t)(L_DMAX + d) * nb + b]);
        s += w;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[d] = s;
}

// flags[t] = 1 if row tile t (128 list positions) holds at least one OOV id: one warp per tile
__global__ void lsh_flags_kernel(const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n, int64_t n_old, uint8_t* __restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t * L_BM >= n) return;
    bool any = false;
#pragma unroll
    for (int i = 0; i < L_BM / 32; ++i) {
        const int64_t r = t * L_BM + i * 32 + lane;
        if (r < n) any |= ids[r * ids_stride] >= n_old;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[t] = any ? 1 : 0;
}

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(L_WORKERS * 32) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tc_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (16-bit pairs packed along K, lane = row) comes from tensor memory,
// where the workers put it with tcgen05.st — no shared-memory round trip
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_st_32x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// the fp32 FMA chain of csrc/lsh.cu (f ascending, original planes): the value that decides a bit on every path.
// Fixer version: all loads are issued before the first FMA (one memory latency instead of one per 8 features); a real
// call, so its 64 operand registers do not count against the worker warps' budget.
__device__ __noinline__ float exact_projection_fast(const float* __restrict__ feat, const float* __restrict__ planes, int F, int64_t fr, int b) {
    const float* xr = feat + fr * F;
    const float* pr = planes + (size_t)b * F;
    float a = 0.f;
    if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(pr)) & 15) == 0) {
        float4 xv[L_FCH / 4], pv[L_FCH / 4];
#pragma unroll
        for (int j = 0; j < L_FCH / 4; ++j)
            if (4 * j < F) { xv[j] = __ldg(reinterpret_cast<const float4*>(xr) + j); pv[j] = __ldg(reinterpret_cast<const float4*>(pr) + j); }
#pragma unroll
        for (int j = 0; j < L_FCH / 4; ++j)
            if (4 * j < F) {
                a = fmaf(xv[j].x, pv[j].x, a); a = fmaf(xv[j].y, pv[j].y, a);
                a = fmaf(xv[j].z, pv[j].z, a); a = fmaf(xv[j].w, pv[j].w, a);
            }
        return a;
    }
    for (int f = 0; f < F; ++f) a = fmaf(__ldg(xr + f), __ldg(pr + f), a);
    return a;
}
// F > 32 (two feature chunks): the same chain, 32 features per round of loads
__device__ __noinline__ float exact_projection_wide(const float* __restrict__ feat, const float* __restrict__ planes, int F, int64_t fr, int b) {
    const float* xr = feat + fr * F;
    const float* pr = planes + (size_t)b * F;
    float a = 0.f;
    if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(pr)) & 15) == 0) {
        for (int f0 = 0; f0 < F; f0 += L_FCH) {
            float4 xv[L_FCH / 4], pv[L_FCH / 4];
#pragma unroll
            for (int j = 0; j < L_FCH / 4; ++j)
                if (f0 + 4 * j < F) { xv[j] = __ldg(reinterpret_cast<const float4*>(xr + f0) + j); pv[j] = __ldg(reinterpret_cast<const float4*>(pr + f0) + j); }
#pragma unroll
            for (int j = 0; j < L_FCH / 4; ++j)
                if (f0 + 4 * j < F) {
                    a = fmaf(xv[j].x, pv[j].x, a); a = fmaf(xv[j].y, pv[j].y, a);
                    a = fmaf(xv[j].z, pv[j].z, a); a = fmaf(xv[j].w, pv[j].w, a);
                }
        }
        return a;
    }
    for (int f = 0; f < F; ++f) a = fmaf(__ldg(xr + f), __ldg(pr + f), a);
    return a;
}
// Worker version (degenerate rows only: a full queue, or the caller wants the multi-hot words): few registers
__device__ __forceinline__ float exact_projection(const float* __restrict__ feat, const float* __restrict__ planes, int F, int64_t fr, int b) {
    const float* xr = feat + fr * F;
    const float* pr = planes + (size_t)b * F;
    float a = 0.f;
#pragma unroll 4
    for (int f = 0; f < F; ++f) a = fmaf(__ldg(xr + f), __ldg(pr + f), a);
    return a;
}

// 8 features of one row (thread = (row, slice)).  The workers look two tiles ahead: a tile's ids (-> feature rows) are
// loaded while the tile before the previous one is projected, its feature rows are requested into L2 one tile ahead
// (no registers) and loaded (L2 hits) when the tile is staged — no chain of dependent global loads between two tiles.
struct Gather {
    float x[8];
    int64_t fr;        // feature row of the list row (-1: not hashed — in-vocab, past the end, or out of range)
};

// feature row of list position (tile, r): -1 = not hashed (in-vocab id, past the end of the list, out of range)
__device__ __forceinline__ int64_t gather_fr(const LshParams& p, int64_t tile, int r) {
    const int64_t rr = tile * L_BM + r;
    int64_t fr = -1;
    if (rr < p.n) {
        const int64_t id = p.ids[rr * p.ids_stride];
        if (id >= p.n_old) {
            fr = feature_row(id, p.prime_pad);
            if (fr < 0 || fr >= p.n_feat_rows) fr = -1;               // out-of-range ids hash nothing (caller bug)
        }
    }
    return fr;
}
__device__ __forceinline__ void gather_load(const LshParams& p, int64_t fr, int f0, Gather& gth) {
    gth.fr = fr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int f = f0 + j;
        gth.x[j] = (fr >= 0 && f < p.F) ? __ldg(p.feat + fr * p.F + f) : 0.f;
    }
}

// in-vocab-only tile: plain gather (bpr.py:111-112), 4 threads per row
__device__ __forceinline__ void copy_iv_tile(const LshParams& p, int64_t tile, int r, int part) {
    const int64_t rr = tile * L_BM + r;
    if (rr >= p.n) return;
    const int64_t id = p.ids[rr * p.ids_stride];
    if (id >= 0 && id < p.n_old && p.iv_table != nullptr)
        for (int d = part * 16; d < p.D && d < part * 16 + 16; ++d)
            store_elem(p.out, p.out_dtype, rr * p.out_stride + d, load_elem(p.iv_table, p.iv_dtype, id * (int64_t)p.D + d));
    if (p.bits_out != nullptr)
        for (int w = part; w < p.words; w += 4) p.bits_out[rr * p.words + w] = 0u;
}

// Everything that is not the common case of a 32-column chunk, out of line and warp-uniform (all 32 lanes call it; the
// projections are read again from tensor memory, 8 columns at a time): near-zero projections are queued for the fixer
// warp (or settled here when the queue is full / the caller wants exact bits), planes >= B are forced to +1.
// Returns the S' bits to force to 1 (x) / to 0 (y), and the number of |R| < tie_eps events seen (z).
struct RareArgs {
    uint32_t acc_addr;            // TMEM address of column 0 of the chunk (this warp's lane quarter)
    uint32_t valid;               // columns < B
    float near; int force, my_oov, exact_inline;
    int b0, row, par;
    int64_t out_row;              // list position of the row (bits_out), -1: past the end
    int64_t fr;                   // feature row
    uint32_t* s_qn; uint32_t* q_ent;
    const float* feat; const float* planes; int F; float tie_eps;
    uint32_t* bits_out; int words;
};
__device__ __noinline__ uint3 lsh_rare_chunk(const RareArgs a) {
    uint32_t setw = 0u, clrw = 0u, ties = 0u, word = 0u;
    const bool any_force = __any_sync(0xffffffffu, a.force && a.my_oov);      // a row with Inf / NaN features
#pragma unroll 1
    for (int gi = 0; gi < 4; ++gi) {
        const uint32_t vg = (a.valid >> (8 * gi)) & 0xffu;                   // warp-uniform
        if (vg == 0u) continue;
        uint32_t u[8];
        tc_ld_32x8(a.acc_addr + (uint32_t)(8 * gi), u);
        tc_wait_ld();
        uint32_t nw = 0u, pw = 0u, wd = 0u;                                  // near zero / sign bit clear (the bit S' carries) / !(R < 0)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            nw |= (fabsf(__uint_as_float(u[j])) >= a.near ? 0u : 1u) << j;    // NaN counts as near
            pw |= ((u[j] >> 31) ^ 1u) << j;
            wd |= (__uint_as_float(u[j]) < 0.f ? 0u : 1u) << j;
        }
        if (a.force) nw = 0xffu;
        nw &= a.my_oov ? vg : 0u;
        (void)any_force;
        if (a.exact_inline) {
            // exact bits now: -0 has the sign bit but is not < 0 (torch_hash.py:57-59), near-zero projections are redone in
            // the fp32 FMA order of csrc/lsh.cu
            while (nw) {
                const int j = __ffs(nw) - 1;
                nw &= nw - 1;
                const float r = exact_projection(a.feat, a.planes, a.F, a.fr, a.b0 + 8 * gi + j);
                wd = (wd & ~(1u << j)) | ((r < 0.f ? 0u : 1u) << j);
                if (fabsf(r) < a.tie_eps) ++ties;
            }
            wd &= a.my_oov ? vg : 0u;
            word |= wd << (8 * gi);
            setw |= wd << (8 * gi); clrw |= (~wd & vg) << (8 * gi);
        } else {
            while (nw) {
                const int j = __ffs(nw) - 1;
                nw &= nw - 1;
                const uint32_t qs = atomicAdd(a.s_qn, 1u);
                if (qs < (uint32_t)L_QCAP) {
                    a.q_ent[qs] = ((uint32_t)a.row << 24) | ((uint32_t)(a.b0 + 8 * gi + j) << 1) | ((pw >> j) & 1u);
                } else {                                                     // queue full (degenerate rows): settle it here
                    const float r = exact_projection(a.feat, a.planes, a.F, a.fr, a.b0 + 8 * gi + j);
                    if (r < 0.f) clrw |= 1u << (8 * gi + j); else setw |= 1u << (8 * gi + j);
                    if (fabsf(r) < a.tie_eps) ++ties;
                }
            }
        }
    }
    if (a.exact_inline && a.out_row >= 0 && (a.b0 >> 5) < a.words) a.bits_out[a.out_row * a.words + (a.b0 >> 5)] = word;
    setw |= ~a.valid;                                                        // planes >= B meet zero bucket rows: +1, taken out of the count by the caller
    return make_uint3(setw, clrw, ties);
}


template <int FC>      // feature chunks of 32 (compile-time: F <= 32 keeps the one-chunk instruction stream)
__global__ void __launch_bounds__(L_THREADS, 1)
tc_lsh_embed_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmW, const LshParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sB = smem;                                        // stages of [128 planes x 128 B] B' tiles
    unsigned char* sW = sB + L_BSTAGES * L_BT_BYTES;                 // stages of 2 x [64 d-rows x 128 B] bucket-table tiles
    unsigned char* tail = sW + L_WSTAGES * L_WT_BYTES;
    float* s_mx = reinterpret_cast<float*>(tail);                    // [4][128] largest |x_i| of the NEXT tile's row slices
    float* s_n2 = s_mx + 4 * L_BM;                                   // [4][128] squared-norm shares of the same
    float* s_cnt = s_n2 + 4 * L_BM;                                  // [4][128] sum of S' over this tile's column slots
    int64_t* s_fr = reinterpret_cast<int64_t*>(s_cnt + 4 * L_BM);    // [2][128] feature row of every tile row (by tile parity)
    uint32_t* q_ent = reinterpret_cast<uint32_t*>(s_fr + 2 * L_BM);  // [2][QCAP] near-zero projections: row << 24 | plane << 1 | bit
    uint32_t* f_ent = q_ent + 2 * L_QCAP;                            // [2][QCAP] signs that differ: row << 24 | plane << 1 | exact bit
    float* swsum = reinterpret_cast<float*>(f_ent + 2 * L_QCAP);     // [64] column sums of the bucket table
    uint32_t* s_qn = reinterpret_cast<uint32_t*>(swsum + L_DMAX);    // [2] queue lengths, [2] flip-list lengths
    uint32_t* s_fn = s_qn + 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_fn + 2);
    uint64_t* a_full = bars;            uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;        uint64_t* b_empty = b_full + L_BSTAGES;
    uint64_t* w_full = b_empty + L_BSTAGES;  uint64_t* w_empty = w_full + L_WSTAGES;
    uint64_t* acc1_full = w_empty + L_WSTAGES;  uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* h_full = acc1_empty + 2;  uint64_t* h_empty = h_full + 2;
    uint64_t* acc2_full = h_empty + 2;  uint64_t* acc2_empty = acc2_full + 1;
    uint64_t* q_full = acc2_empty + 1;  uint64_t* f_full = q_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(f_full + 2);

    // Physical warps 0-15 are the workers (warp & 3 = TMEM lane quarter), 16-19 the single-purpose roles: the
    // sub-partition schedulers favour the highest warp id (B300_MICROARCH.md), so a role warp that becomes ready issues
    // ahead of the workers of its scheduler.
    const int warp = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + L_BM - 1) / L_BM;
    const int NT = p.NT;
    const int SB = p.res_b ? NT * FC : L_BSTAGES;                  // stages in use (one B' tile per N tile and feature chunk)
    const int NW = NT * p.wsplit;                                    // bucket-table tiles per row tile
    const int SW = p.res_w ? NW : L_WSTAGES;
    // next row tile of this CTA (at or after t) that holds an OOV id — one byte per tile, written by lsh_flags_kernel
    auto next_oov = [&](int64_t t) { while (t < n_tiles && p.flags[t] == 0) t += gridDim.x; return t; };

    if (warp == W_TMA && lane == 0) { tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmW); }
    if (warp == W_MMA1 && lane == 0) {
        mbar_init(a_full, L_WORKERS); mbar_init(a_empty, 1);
        for (int s = 0; s < L_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < L_WSTAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc1_full[a], 1); mbar_init(&acc1_empty[a], L_WORKERS);
            mbar_init(&h_full[a], L_WORKERS); mbar_init(&h_empty[a], 1);
            mbar_init(&q_full[a], 1); mbar_init(&f_full[a], 1);
        }
        mbar_init(acc2_full, 1); mbar_init(acc2_empty, L_WORKERS);
        s_qn[0] = s_qn[1] = s_fn[0] = s_fn[1] = 0u;
        fence_barrier_init();
    }
    if (warp == W_FIXER) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < L_DMAX) swsum[threadIdx.x] = p.wsum[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t ACC2_COL = 256;                               // 64 fp32 columns: S' W
    constexpr uint32_t H_COL = 320;                                  // 2 x 64 columns: S' [128 x 128] fp16, two per column
    constexpr uint32_t A_COL = 448;                                  // 32 columns per feature chunk: A' [128 x 64] fp16, two per column (448-511)

    // The single-thread roles run with ALL 32 lanes of their warp in uniform control flow and only predicate the TMA / MMA /
    // commit instructions on one elected lane.  Under a divergent `if (lane == 0)` ptxas cannot keep descriptors and
    // addresses in uniform registers and wraps every UTCHMMA / UTMALDG in a vote loop (ELECT + 3 R2UR.BROADCAST +
    // BRA.U.ANY, ~14 dependent instructions): 125-200 cycles per MMA next to four busy worker warps on the same
    // scheduler (trace: scripts/trace_lsh.py; the tensor pipe itself keeps its nominal rate: scripts/ubench/mma_rate.cu).
    const bool leader = elect_one();
    if (warp == W_TMA) {
        // ===================== TMA: B' tiles and transposed bucket-table tiles (one ring each, filled in consumption order) =====================
        int bst = 0, wst = 0; uint32_t bphase = 0, wphase = 0;
        bool b_done = false, w_done = false;                          // resident operands are loaded once
        // Side job: the fp32 -> bf16 cast of the in-vocab rows of the same item table (bpr.py:111-112 for a contiguous id
        // range).  As its own launch this memory-bound copy cannot share an SM with the 640-thread / 96-register LSH CTAs
        // and runs in front of them; here this warp does it a 4 KB group at a time WHILE IT WAITS for a free stage of the
        // bucket-table ring (the ring is full then, so the consumers are not held up), in the shadow of the latency-bound
        // GEMM pipeline (a few percent of the DRAM bandwidth).  A group = 32 lanes x 4 x 8 elements; the lines of the group
        // four ahead are pulled into L2 first.
        const int64_t cast_groups = (p.cast_n8 + 127) >> 7;
        int64_t cast_g = blockIdx.x;
        auto cast_step = [&]() {
            const float4* src = reinterpret_cast<const float4*>(p.cast_src);
            uint4* dst = reinterpret_cast<uint4*>(p.cast_dst);
            const int64_t g = cast_g, gp = cast_g + 4 * (int64_t)gridDim.x;
            cast_g += gridDim.x;
            if (gp < cast_groups && (gp << 7) + lane * 4 < p.cast_n8)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(src) + (gp << 12) + lane * 128));
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t i8 = (g << 7) + j * 32 + lane;
                if (i8 < p.cast_n8) { v[2 * j] = __ldcs(src + 2 * i8); v[2 * j + 1] = __ldcs(src + 2 * i8 + 1); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t i8 = (g << 7) + j * 32 + lane;
                if (i8 < p.cast_n8) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * j].x, v[2 * j].y), b = __floats2bfloat162_rn(v[2 * j].z, v[2 * j].w);
                    __nv_bfloat162 c = __floats2bfloat162_rn(v[2 * j + 1].x, v[2 * j + 1].y), d = __floats2bfloat162_rn(v[2 * j + 1].z, v[2 * j + 1].w);
                    dst[i8] = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                                         *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
                }
            }
        };
        for (int64_t t = next_oov(blockIdx.x); t < n_tiles && !(b_done && w_done); t = next_oov(t + gridDim.x)) {
            for (int nt = 0; nt < NT; ++nt) {
                if (!b_done)
                    for (int c = 0; c < FC; ++c) {                  // one B' tile per feature chunk, in consumption order
                        if (!p.res_b) mbar_wait_spin(&b_empty[bst], bphase ^ 1);
                        if (leader) {
                            mbar_arrive_expect_tx(&b_full[bst], L_BT_BYTES);
                            tma_load_2d(sB + bst * L_BT_BYTES, &tmB, &b_full[bst], 0, (nt * FC + c) * L_BN);
                        }
                        __syncwarp();
                        if (++bst == SB) { bst = 0; bphase ^= 1; }
                    }
                if (!w_done)
                    for (int pc = 0; pc < p.wsplit; ++pc) {
                        if (!p.res_w) {
                            while (cast_g < cast_groups && !mbar_test_wait(&w_empty[wst], wphase ^ 1)) cast_step();
                            mbar_wait_spin(&w_empty[wst], wphase ^ 1);
                        }
                        if (leader) {
                            mbar_arrive_expect_tx(&w_full[wst], L_WT_BYTES);
                            tma_load_2d(sW + wst * L_WT_BYTES, &tmW, &w_full[wst], nt * L_BN, pc * L_DMAX);
                            tma_load_2d(sW + wst * L_WT_BYTES + L_WT_BYTES / 2, &tmW, &w_full[wst], nt * L_BN + 64, pc * L_DMAX);
                        }
                        __syncwarp();
                        if (++wst == SW) { wst = 0; wphase ^= 1; }
                    }
            }
            b_done = p.res_b != 0;
            w_done = p.res_w != 0;
        }
        while (cast_g < cast_groups) cast_step();                    // what is left of the side job
    } else if (warp == W_MMA1) {
        // ===================== MMA issuer 1: projections =====================
        constexpr uint32_t idesc1 = make_idesc_bf16_f32(L_BM, L_BN) & ~((7u << 7) | (7u << 10));   // A, B = fp16 (format 0)
        int bs = 0; uint32_t bph = 0;
        int64_t g1 = 0;                    // N tiles issued since kernel start
        int64_t T = 0;                     // row tiles with OOV ids done by this CTA
        int trace_n = 0;
#define trace_tag g1
        for (int64_t t = next_oov(blockIdx.x); t < n_tiles; t = next_oov(t + gridDim.x)) {
            if (leader) LTRACE(0, 0);
            mbar_wait_spin(a_full, (uint32_t)(T & 1));
            for (int nt = 0; nt < NT; ++nt, ++g1) {
                const int buf = (int)(g1 & 1);
                if (leader) LTRACE(0, 1);
                mbar_wait_spin(&acc1_empty[buf], (uint32_t)(((g1 >> 1) & 1) ^ 1));
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * L_BN);
                for (int c = 0; c < FC; ++c) {                      // K loop over 32-feature chunks (F <= 32: one round)
                    mbar_wait_spin(&b_full[bs], p.res_b ? 0u : bph);  // resident: phase 0 completed once and for all
                    tc_fence_after();
                    const uint32_t a_tmem = tmem_base + A_COL + (uint32_t)(c * 32);
                    const uint64_t bdesc = make_sw128_desc(smem_u32(sB + bs * L_BT_BYTES));
                    if (leader) {
                        LTRACE(0, 2);
                        // K = 16 fp16 = 8 TMEM columns of A' (x0: columns 0-15, x1: 16-31) / 32 B of a B' row (p0: bytes 0-63, p1: 64-127)
                        tc_mma_f16_ts(d_tmem, a_tmem + 0, bdesc + 0, idesc1, c ? 1u : 0u);    // x0 p0
                        tc_mma_f16_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc1, 1u);
                        tc_mma_f16_ts(d_tmem, a_tmem + 16, bdesc + 0, idesc1, 1u);   // x1 p0
                        tc_mma_f16_ts(d_tmem, a_tmem + 24, bdesc + 2, idesc1, 1u);
                        tc_mma_f16_ts(d_tmem, a_tmem + 0, bdesc + 4, idesc1, 1u);    // x0 p1
                        tc_mma_f16_ts(d_tmem, a_tmem + 8, bdesc + 6, idesc1, 1u);
                        if (!p.res_b) tc_commit(&b_empty[bs]);
                        if (c == FC - 1) {
                            tc_commit(&acc1_full[buf]);
                            if (nt == NT - 1) tc_commit(a_empty);     // A' may be rebuilt for the next row tile
                        }
                        LTRACE(0, 3);
                    }
                    __syncwarp();
                    if (++bs == SB) { bs = 0; bph ^= 1; }
                }
            }
            ++T;
        }
#undef trace_tag
    } else if (warp == W_MMA2) {
        // ===================== MMA issuer 2: acc2 += S' W, S' read from TMEM =====================
        constexpr uint32_t idesc2 = make_idesc_bf16_f32(L_BM, L_DMAX) & ~((7u << 7) | (7u << 10));   // A, B = fp16 (format 0)
        int ws = 0; uint32_t wph = 0;
        int64_t g2 = 0;
        int64_t T = 0;
        int trace_n = 0;
#define trace_tag g2
        for (int64_t t = next_oov(blockIdx.x); t < n_tiles; t = next_oov(t + gridDim.x)) {
            const uint32_t d_tmem = tmem_base + ACC2_COL;
            for (int j = 0; j < NT; ++j, ++g2) {
                const int hb = (int)(g2 & 1);
                if (leader) LTRACE(1, 4);
                mbar_wait_spin(&h_full[hb], (uint32_t)((g2 >> 1) & 1));
                if (j == 0) mbar_wait_spin(acc2_empty, (uint32_t)((T & 1) ^ 1));
                if (leader) LTRACE(1, 5);
                for (int pc = 0; pc < p.wsplit; ++pc) {
                    mbar_wait_spin(&w_full[ws], p.res_w ? 0u : wph);
                    tc_fence_after();
                    const uint32_t a_tmem = tmem_base + H_COL + (uint32_t)(hb * 64);
                    const uint64_t bdesc0 = make_sw128_desc(smem_u32(sW + ws * L_WT_BYTES));
                    const uint64_t bdesc1 = make_sw128_desc(smem_u32(sW + ws * L_WT_BYTES + L_WT_BYTES / 2));
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)                   // K = 16 fp16 = 8 TMEM columns / 32 B of smem
                            tc_mma_f16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), (k < 4 ? bdesc0 + (uint64_t)(2 * k) : bdesc1 + (uint64_t)(2 * (k - 4))),
                                          idesc2, (j | pc | k) ? 1u : 0u);
                        if (!p.res_w) tc_commit(&w_empty[ws]);
                    }
                    __syncwarp();
                    if (++ws == SW) { ws = 0; wph ^= 1; }
                }
                if (leader) {
                    tc_commit(&h_empty[hb]);
                    if (j == NT - 1) tc_commit(acc2_full);
                    LTRACE(1, 6);
                }
                __syncwarp();
            }
            ++T;
        }
#undef trace_tag
    } else if (warp == W_FIXER) {
        // ===================== fixer: settles the queued near-zero projections of a row tile while the workers go on =====================
        unsigned int ties = 0;
        int64_t T = 0;
        for (int64_t t = next_oov(blockIdx.x); t < n_tiles; t = next_oov(t + gridDim.x), ++T) {
            const int par = (int)(T & 1);
            mbar_wait(&q_full[par], (uint32_t)((T >> 1) & 1));                  // the queue of this tile is complete
            const uint32_t nq = min(s_qn[par], (uint32_t)L_QCAP);
            for (uint32_t e = (uint32_t)lane; e < nq; e += 32) {
                const uint32_t ent = q_ent[par * L_QCAP + e];
                const int r = (int)(ent >> 24), b = (int)((ent >> 1) & 0x7fffffu);
                const float a = FC == 1 ? exact_projection_fast(p.feat, p.planes, p.F, s_fr[par * L_BM + r], b)
                                        : exact_projection_wide(p.feat, p.planes, p.F, s_fr[par * L_BM + r], b);
                const uint32_t bit = a < 0.f ? 0u : 1u;
                if (fabsf(a) < p.tie_eps) ++ties;
                if (bit != (ent & 1u)) f_ent[par * L_QCAP + atomicAdd(&s_fn[par], 1u)] = (ent & ~1u) | bit;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&f_full[par]);                           // the flip list is complete
        }
        if (p.tie_count != nullptr) {
            for (int o = 16; o; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
            if (lane == 0 && ties) atomicAdd(p.tie_count, (unsigned long long)ties);
        }
    } else {
#define trace_tag g
        // ===================== workers =====================
        const int wk = warp;
        const int q = wk & 3;                      // TMEM lane quarter
        const int slot = wk >> 2;                  // column quarter of every N tile / 8-feature slice / count slot / 16-column output slice
        const int row = q * 32 + lane;             // row of the tile this thread owns
        const int wtid = wk * 32 + lane;           // 0..511
        const int gr = wtid >> 2, gpart = wtid & 3;   // in-vocab copy role: row, 16-column slice
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a_lane = lane_base + A_COL;
        unsigned int my_ties = 0;
        int g = 0, T = 0;
        int trace_n = 0;
        const int trole = (lane == 0 && wk == 0) ? 2 : ((lane == 0 && wk == OOV_TRACE_W2) ? 3 : -1);
#ifdef OOV_LSH_TRACE
#define WTRACE(ev) do { if (trole >= 0) LTRACE(trole, ev); } while (0)
#else
#define WTRACE(ev) do { (void)trole; } while (0)
#endif
        const float pn_min = __uint_as_float(*reinterpret_cast<const unsigned int*>(p.pn_min));
        // fp16 pair: sign bits / (+1, +1).  Kept in registers (opaque to the compiler) so that (x & SIGNS) ^ ONES is ONE
        // three-register LOP3 instead of two with an immediate each: the alu pipe (PRMT / LOP3 / FMNMX, one warp
        // instruction per two cycles per scheduler) is what bounds the workers.
        uint32_t SIGNS, ONES;
        asm volatile("mov.b32 %0, 0x80008000;" : "=r"(SIGNS));
        asm volatile("mov.b32 %0, 0x3C003C00;" : "=r"(ONES));
        const bool exact_inline = p.bits_out != nullptr;              // the caller wants the multi-hot words: no deferral

        // The row state of the tile being projected (set by stage_tile from the prefetched features)
        bool my_oov = false;
        float near = 0.f;
        bool force = false;

        // Turn the prefetched features of tile `Tn` (its parity selects the s_fr bank) into A' = [x0 | x1] in TMEM and the
        // row's near-zero threshold.  Called by every worker thread (it holds a worker barrier).
        auto stage_tile = [&](int64_t fr, int Tn, bool wait_a_empty) {
            Gather gth;
            gather_load(p, fr, slot * 8, gth);
            float mx = 0.f, n2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { mx = fmaxf(mx, fabsf(gth.x[j])); n2 = fmaf(gth.x[j], gth.x[j], n2); }
            bool bad2 = false;
            if (FC > 1) {
                // second feature chunk (F > 32): only its maximum and squared norm are needed before the row scale is known;
                // the values are loaded again (L1 / L2 hits) when they are converted — no registers held across the barrier
                Gather g2;
                gather_load(p, fr, L_FCH + slot * 8, g2);
#pragma unroll
                for (int j = 0; j < 8; ++j) { mx = fmaxf(mx, fabsf(g2.x[j])); n2 = fmaf(g2.x[j], g2.x[j], n2); bad2 |= !(fabsf(g2.x[j]) < INFINITY); }
            }
            bool bad = bad2;
#pragma unroll
            for (int j = 0; j < 8; ++j) bad |= !(fabsf(gth.x[j]) < INFINITY);      // Inf / NaN features: every bit is re-evaluated
            s_mx[slot * L_BM + row] = bad ? INFINITY : mx;
            s_n2[slot * L_BM + row] = n2;
            if (slot == 0) s_fr[(Tn & 1) * L_BM + row] = gth.fr;
            WTRACE(17);
            worker_bar();
            WTRACE(18);                                              // also: the previous tile's queue and counts are complete
            if (wtid == 0) {
                s_qn[Tn & 1] = 0u; s_fn[Tn & 1] = 0u;                  // the bank of tile Tn (last used by Tn - 2, long settled)
                if (wait_a_empty) mbar_arrive(&q_full[(Tn - 1) & 1]);  // the fixer may settle tile Tn - 1 now
            }
            const float m4 = fmaxf(fmaxf(s_mx[row], s_mx[L_BM + row]), fmaxf(s_mx[2 * L_BM + row], s_mx[3 * L_BM + row]));
            const float nn = (s_n2[row] + s_n2[L_BM + row]) + (s_n2[2 * L_BM + row] + s_n2[3 * L_BM + row]);
            // power-of-two row scale: largest |x_i| -> [2^13, 2^14) (exact; the sign of the projection does not change)
            const uint32_t e = (__float_as_uint(m4) >> 23) & 0xffu;
            const float sc = __uint_as_float((e >= 13u ? (e <= 254u ? 267u - e : 1u) : 254u) << 23);
            // |sc x|: from the squared norm unless that may have under- / overflowed, then from sqrt(F) max|x_i| (looser)
            const float xn = ((m4 > 1e-15f) & (m4 < 1e15f)) ? sqrtf(nn) * sc : (FC > 1 ? 8.f : 5.6568542f) * m4 * sc;
            my_oov = gth.fr >= 0;
            force = !(m4 < INFINITY) || !(xn < INFINITY);
            // |R' - sc x.p^| < L_NEAR_REL |sc x|; |x.p| < tie_eps (the reported ties) lies inside 256 sc tie_eps / |p|
            near = fmaxf(L_NEAR_REL * xn, 2.f * L_PSCALE * p.tie_eps * sc / pn_min);
            uint32_t c0[4], c1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = gth.x[2 * j] * sc, b = gth.x[2 * j + 1] * sc;
                const __half a0 = __float2half_rn(a), b0 = __float2half_rn(b);
                c0[j] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(b0) << 16);
                c1[j] = pack_f16x2(a - __half2float(a0), b - __half2float(b0));
            }
            if (wait_a_empty) {
                mbar_wait(a_empty, (uint32_t)((Tn - 1) & 1));          // the previous tile's GEMM1s have read A'
                tc_fence_after();
            }
            WTRACE(19);
            // K element k lives in column k / 2: the 8 features are 4 columns at offset 4 * slot of each 16-column piece
            tc_st_32x4(a_lane + 0 * 16 + slot * 4, c0[0], c0[1], c0[2], c0[3]);
            tc_st_32x4(a_lane + 1 * 16 + slot * 4, c1[0], c1[1], c1[2], c1[3]);
            if (FC > 1) {                                           // chunk 1 of A': columns 32-63, same layout
                gather_load(p, gth.fr, L_FCH + slot * 8, gth);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = gth.x[2 * j] * sc, b = gth.x[2 * j + 1] * sc;
                    const __half a0 = __float2half_rn(a), b0 = __float2half_rn(b);
                    c0[j] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(b0) << 16);
                    c1[j] = pack_f16x2(a - __half2float(a0), b - __half2float(b0));
                }
                tc_st_32x4(a_lane + 32 + 0 * 16 + slot * 4, c0[0], c0[1], c0[2], c0[3]);
                tc_st_32x4(a_lane + 32 + 1 * 16 + slot * 4, c1[0], c1[1], c1[2], c1[3]);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        };

        // first two tiles with OOV ids (in-vocab-only tiles on the way are plain copies); later ones are found while projecting
        int64_t t = blockIdx.x;
        while (t < n_tiles && p.flags[t] == 0) { copy_iv_tile(p, t, gr, gpart); t += gridDim.x; }
        int64_t tn = t + gridDim.x;
        while (tn < n_tiles && p.flags[tn] == 0) { copy_iv_tile(p, tn, gr, gpart); tn += gridDim.x; }
        int64_t fr_n = tn < n_tiles ? gather_fr(p, tn, row) : -1;     // feature row of this thread's row in tile tn
        if (t < n_tiles) stage_tile(gather_fr(p, t, row), 0, false);  // a_empty: nothing has read A' yet
        while (t < n_tiles) {
            const int64_t row0 = t * L_BM;
            const int par = (int)(T & 1);
            const bool cur_oov = my_oov;                              // (stage_tile moves my_oov / near / force on to the next tile)
            // look ahead: request tile tn's feature rows into L2, load the flag and the ids of the tile after it (consumed at
            // the end of this tile: the loads fly while this tile is projected)
            if (fr_n >= 0 && slot * 8 < p.F) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.feat + fr_n * p.F + slot * 8));
            if (FC > 1 && fr_n >= 0 && L_FCH + slot * 8 < p.F) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.feat + fr_n * p.F + L_FCH + slot * 8));
            int64_t tnn = tn + gridDim.x;
            const uint8_t flag_nn = tnn < n_tiles ? p.flags[tnn] : (uint8_t)1;
            int64_t fr_nn = tnn < n_tiles ? gather_fr(p, tnn, row) : -1;

            __half2 cn0 = __float2half2_rn(0.f), cn1 = cn0;           // sum of this thread's S' words (two chains)
            int padc = 0;                                             // planes >= B among them (forced to +1)
            // ---- per N tile: projections -> signs -> S'
            for (int nt = 0; nt < NT; ++nt, ++g) {
                const int buf = g & 1;
                const uint32_t bpar = (uint32_t)((g >> 1) & 1);
                const uint32_t acc_addr = lane_base + (uint32_t)(buf * L_BN + slot * 32);
                WTRACE(7);
                mbar_wait(&acc1_full[buf], bpar);
                tc_fence_after();
                WTRACE(8);
                // asked now, needed before the store below: the barrier round trip hides under the load and the conversion
                const bool h_ok = mbar_try_wait(&h_empty[buf], bpar ^ 1);
                // S' pair = (+1, +1) with the sign bits of the two projections (2 instructions per pair) and the smallest |R|
                // of every 8-column group (one 3-input min per pair); 16 columns at a time (registers)
                uint32_t hw[16];
                float gm[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    uint32_t v[16];
                    tc_ld_32x16(acc_addr + (uint32_t)(16 * qq), v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(hw[8 * qq + i]) : "r"(__byte_perm(v[2 * i], v[2 * i + 1], 0x7030)), "r"(SIGNS), "r"(ONES));
                        gm[2 * qq + (i >> 2)] = fminf(gm[2 * qq + (i >> 2)], fminf(fabsf(__uint_as_float(v[2 * i])), fabsf(__uint_as_float(v[2 * i + 1]))));
                    }
                }
                WTRACE(9);
                const float mn = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
                const int b0 = nt * L_BN + slot * 32;                 // plane of column 0
                // rare (warp-uniform: the out-of-line path reads tensor memory with warp-collective loads): a projection
                // too close to zero to trust its sign, a row with Inf / NaN features, planes >= B in the chunk (their
                // projections are exact zeros), or the caller wants the multi-hot words
#ifdef OOV_LSH_NORARE   /* timing experiment only: results are wrong without this path */
                if (false) {
#else
                if ((b0 + 32 > p.B) | exact_inline | __any_sync(0xffffffffu, ((mn < near) | force) & my_oov)) {
#endif
                    RareArgs ra;
                    ra.acc_addr = acc_addr;
                    ra.valid = (b0 + 32 <= p.B) ? 0xffffffffu : ((b0 >= p.B) ? 0u : ((1u << (p.B - b0)) - 1u));
                    ra.near = near; ra.force = force; ra.my_oov = my_oov; ra.exact_inline = exact_inline;
                    ra.b0 = b0; ra.row = row; ra.par = par;
                    ra.out_row = row0 + row < p.n ? row0 + row : -1;
                    ra.fr = s_fr[par * L_BM + row];
                    ra.s_qn = &s_qn[par]; ra.q_ent = q_ent + par * L_QCAP;
                    ra.feat = p.feat; ra.planes = p.planes; ra.F = p.F; ra.tie_eps = p.tie_eps;
                    ra.bits_out = p.bits_out; ra.words = p.words;
                    const uint3 fx = lsh_rare_chunk(ra);
                    my_ties += fx.z;
                    if (fx.x | fx.y) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const uint32_t s2 = (fx.x >> (2 * i)) & 3u, c2 = (fx.y >> (2 * i)) & 3u;
                            uint32_t w = hw[i];
                            w &= ~(((s2 & 1u) ? 0x8000u : 0u) | ((s2 & 2u) ? 0x80000000u : 0u));
                            w |= ((c2 & 1u) ? 0x8000u : 0u) | ((c2 & 2u) ? 0x80000000u : 0u);
                            hw[i] = w;
                        }
                    }
                    padc += __popc(~ra.valid);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);         // the projections of this N tile are no longer needed
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    cn0 = __hadd2(cn0, *reinterpret_cast<const __half2*>(&hw[i]));
                    cn1 = __hadd2(cn1, *reinterpret_cast<const __half2*>(&hw[i + 1]));
                }
                WTRACE(10);
                if (!h_ok) mbar_wait(&h_empty[buf], bpar ^ 1);        // GEMM2 of the previous use of this S' buffer is done
                tc_fence_after();
                WTRACE(11);
                tc_st_32x16(lane_base + H_COL + (uint32_t)(buf * 64 + slot * 16), hw);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[buf]);
                WTRACE(12);
            }
            {
                const __half2 c = __hadd2(cn0, cn1);
                s_cnt[slot * L_BM + row] = (__low2float(c) + __high2float(c)) - (float)padc;
            }
            // ---- A' of the next tile (its features arrived long ago), so the tensor core can go on while we finish; the
            //      barrier inside also completes this tile's queue and hands it to the fixer warp
            if (tn < n_tiles) stage_tile(fr_n, T + 1, true);
            else {
                worker_bar();
                if (wtid == 0) mbar_arrive(&q_full[par]);
            }
            WTRACE(13);
            // ---- final: out = (S' W + colsum W + corrections) / (sum S' + B)   [= 2 H W / 2 count]
            mbar_wait(acc2_full, (uint32_t)(T & 1));
            tc_fence_after();
            WTRACE(15);
            uint32_t a[16];
            tc_ld_32x16(lane_base + ACC2_COL + (uint32_t)(slot * 16), a);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);                   // GEMM2 of the next tile may start
            mbar_wait(&f_full[par], (uint32_t)((T >> 1) & 1));        // the fixer's list of signs that differ
            WTRACE(14);
            const int64_t r = row0 + row;
            if (r < p.n) {
                const int d0 = slot * 16;
                const size_t osz = p.out_dtype == OOV_F32 ? 4 : 2;
                char* orow = reinterpret_cast<char*>(p.out) + (size_t)r * p.out_stride * osz;
                if (cur_oov) {
                    float num[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) num[i] = __uint_as_float(a[i]) + swsum[d0 + i];
                    float den = ((s_cnt[row] + s_cnt[L_BM + row]) + (s_cnt[2 * L_BM + row] + s_cnt[3 * L_BM + row])) + (float)p.B;   // 2 x count, exact
                    const uint32_t nf = s_fn[par];
                    for (uint32_t e = 0; e < nf; ++e) {
                        const uint32_t ent = f_ent[par * L_QCAP + e];
                        if ((int)(ent >> 24) != row) continue;
                        const int64_t b = (int64_t)((ent >> 1) & 0x7fffffu);
                        const float sg = (ent & 1u) ? 2.f : -2.f;     // S' goes from -1 to +1 or back
                        den += sg;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float w = __half2float(p.Wt[(size_t)(d0 + i) * p.nb + b]);
                            if (p.wsplit == 2) w += __half2float(p.Wt[(size_t)(L_DMAX + d0 + i) * p.nb + b]);
                            num[i] = fmaf(sg, w, num[i]);
                        }
                    }
                    float o[16];
                    const float rden = den == 0.f ? __uint_as_float(0x7FC00000u) : __frcp_rn(den);   // 0 / 0 -> NaN (lsh_embedder.py:158)
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = num[i] * rden;
                    if (p.out_dtype == OOV_BF16 && d0 + 16 <= p.D && ((reinterpret_cast<uintptr_t>(orow) + d0 * 2) & 15) == 0) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(o[2 * i], o[2 * i + 1]);
                        uint4* op = reinterpret_cast<uint4*>(orow + d0 * 2);
                        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (d0 + i < p.D) store_elem(p.out, p.out_dtype, r * p.out_stride + d0 + i, o[i]);
                    }
                } else if (p.iv_table != nullptr) {
                    const int64_t id = p.ids[r * p.ids_stride];       // in-vocab gather (bpr.py:111-112)
                    if (id >= 0 && id < p.n_old)
                        for (int i = 0; i < 16; ++i)
                            if (d0 + i < p.D)
                                store_elem(p.out, p.out_dtype, r * p.out_stride + d0 + i, load_elem(p.iv_table, p.iv_dtype, id * (int64_t)p.D + d0 + i));
                }
            }
            if (flag_nn == 0) {                                       // rare: in-vocab-only tiles ahead — find the next OOV tile the slow way
                while (tnn < n_tiles && p.flags[tnn] == 0) { copy_iv_tile(p, tnn, gr, gpart); tnn += gridDim.x; }
                fr_nn = tnn < n_tiles ? gather_fr(p, tnn, row) : -1;
            }
            WTRACE(16);
            ++T;
            t = tn; tn = tnn; fr_n = fr_nn;
        }
        if (p.tie_count != nullptr) {
            for (int o = 16; o; o >>= 1) my_ties += __shfl_xor_sync(0xffffffffu, my_ties, o);
            if (lane == 0 && my_ties) atomicAdd(p.tie_count, (unsigned long long)my_ties);
        }
#undef trace_tag
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_FIXER) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// the side cast as its own launch (operands not resident: the TMA warp stays busy; or no OOV tile at all)
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 u = __ldcs(src + 2 * i), w = __ldcs(src + 2 * i + 1);
        __nv_bfloat162 a = __floats2bfloat162_rn(u.x, u.y), b = __floats2bfloat162_rn(u.z, u.w);
        __nv_bfloat162 c = __floats2bfloat162_rn(w.x, w.y), d = __floats2bfloat162_rn(w.z, w.w);
        dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                            *reinterpret_cast<uint32_t*>(&d));
    }
}
int launch_cast_f32_bf16(const float* src, void* dst, int64_t n_elems, cudaStream_t st) {
    if (n_elems <= 0) return OOV_OK;
    int64_t blocks = cdiv(n_elems / 8, 256);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), n_elems / 8);
    OOV_LAUNCH_CHECK("cast_f32_bf16_kernel");
    return OOV_OK;
}

// ---------------------------------------------------------------- host
bool lsh_tc_supported(int F, int B, int D) { return F >= 1 && F <= L_FMAX && D >= 1 && D <= L_DMAX && B >= 1 && B <= (1 << 22); }

static size_t lsh_bp_bytes(int B) { return align_up((size_t)(L_FMAX / L_FCH) * cdiv(B, L_BN) * L_BN * 64 * 2, 1024); }   // sized for two feature chunks
static size_t lsh_wt_bytes(int B) { return align_up((size_t)2 * L_DMAX * cdiv(B, L_BN) * L_BN * 2, 1024); }
static size_t lsh_flag_bytes(int64_t n) { return align_up((size_t)cdiv(n, L_BM), 256); }
size_t lsh_tc_workspace(int64_t n, int B) { return lsh_bp_bytes(B) + lsh_wt_bytes(B) + 512 + lsh_flag_bytes(n) + 1024; }

int lsh_tc_run(const float* feat, int64_t n_feat_rows, int F, const float* planes, int B, const void* W, int w_dtype,
               const oov_rows* rows, float tie_eps, uint32_t* bits_out, unsigned long long* tie_count, void* workspace,
               size_t workspace_bytes, cudaStream_t st, const float* cast_src, void* cast_dst, int64_t cast_elems) {
    OOV_REQUIRE(workspace && workspace_bytes >= lsh_tc_workspace(rows->n, B), OOV_ERR_WORKSPACE, "oov_lsh_embed (tcgen05): workspace %zu < %zu",
                workspace_bytes, lsh_tc_workspace(rows->n, B));
    char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __half* Bp = reinterpret_cast<__half*>(ws);
    __half* Wt = reinterpret_cast<__half*>(ws + lsh_bp_bytes(B));
    float* pn_min = reinterpret_cast<float*>(ws + lsh_bp_bytes(B) + lsh_wt_bytes(B));
    float* wsum = pn_min + 16;
    uint8_t* flags = reinterpret_cast<uint8_t*>(ws + lsh_bp_bytes(B) + lsh_wt_bytes(B) + 512);
    cudaError_t ce = cudaMemsetAsync(pn_min, 0x7f, 4, st);           // 0x7f7f7f7f = 3.4e38: "no plane seen"
    OOV_REQUIRE(ce == cudaSuccess, OOV_ERR_CUDA, "cudaMemsetAsync(pn_min): %s", cudaGetErrorString(ce));
    const int NT = (int)cdiv(B, L_BN);
    const int64_t nb = (int64_t)NT * L_BN;
    LshParams p{};
    p.feat = feat; p.n_feat_rows = n_feat_rows; p.F = F; p.FC = (F + L_FCH - 1) / L_FCH; p.planes = planes; p.B = B; p.NT = NT;
    p.ids = rows->ids; p.ids_stride = rows->ids_stride; p.n = rows->n; p.n_old = rows->n_old; p.prime_pad = rows->prime_pad;
    p.iv_table = rows->iv_table; p.iv_dtype = rows->iv_dtype; p.out = rows->out; p.out_dtype = rows->out_dtype;
    p.out_stride = rows->out_stride; p.D = rows->D;
    p.wsplit = rows->out_dtype == OOV_F32 ? 2 : 1;   // fp16 hi (+ lo) pieces of the fp32 bucket table: 2^-12 (2^-23) relative
    p.res_b = NT * p.FC <= L_BSTAGES ? 1 : 0;
    p.res_w = NT * p.wsplit <= L_WSTAGES ? 1 : 0;
    p.tie_eps = tie_eps; p.bits_out = bits_out; p.words = (B + 31) / 32; p.tie_count = tie_count; p.pn_min = pn_min; p.wsum = wsum;
    p.Wt = Wt; p.nb = nb; p.flags = flags;
    const int64_t n_tiles_ = cdiv(rows->n, L_BM);
    static int fuse_cast = -1;
    if (fuse_cast < 0) { const char* e = getenv("OOV_LSH_FUSE_CAST"); fuse_cast = e ? atoi(e) : 1; }   // A/B knob (profiling)
    if (cast_elems > 0) {
        // every CTA's TMA warp takes a share: the grid must fill the machine
        if (fuse_cast && n_tiles_ >= num_sms()) {
            p.cast_src = cast_src; p.cast_dst = reinterpret_cast<__nv_bfloat16*>(cast_dst); p.cast_n8 = cast_elems / 8;
        } else {
            int rc0 = launch_cast_f32_bf16(cast_src, cast_dst, cast_elems, st);
            if (rc0) return rc0;
        }
    }
#ifdef OOV_LSH_TRACE
    if (const char* tp = getenv("OOV_LSH_TRACE_PTR")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0));   // profiling only
#endif

    const int64_t n_tiles = cdiv(rows->n, L_BM);
    lsh_flags_kernel<<<(unsigned)cdiv(n_tiles, 8), 256, 0, st>>>(rows->ids, rows->ids_stride, rows->n, rows->n_old, flags);
    OOV_LAUNCH_CHECK("lsh_flags_kernel");
    const int64_t pack_threads = nb * L_DMAX;                        // >= nb * FC * 32 (FC <= 2)
    lsh_pack_kernel<<<(unsigned)cdiv(pack_threads, 256), 256, 0, st>>>(planes, B, F, NT, W, w_dtype, rows->D, p.wsplit, Bp, Wt, pn_min);
    OOV_LAUNCH_CHECK("lsh_pack_kernel");
    lsh_wsum_kernel<<<L_DMAX, 32, 0, st>>>(Wt, nb, p.wsplit, wsum);
    OOV_LAUNCH_CHECK("lsh_wsum_kernel");

    CUtensorMap tmB, tmW;
    int rc = make_tmap_bf16_2d(&tmB, Bp, 64, (uint64_t)nb * p.FC, 64 * 2, L_BN);                                     // fp16: same 2-byte boxes
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmW, Wt, (uint64_t)nb, (uint64_t)(p.wsplit * L_DMAX), (uint64_t)nb * 2, L_DMAX);
    if (rc) return rc;
    auto kern = p.FC == 1 ? tc_lsh_embed_kernel<1> : tc_lsh_embed_kernel<2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM);
    OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_lsh_embed_kernel): %s", cudaGetErrorString(e));
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    kern<<<grid, L_THREADS, L_SMEM, st>>>(tmB, tmW, p);
    OOV_LAUNCH_CHECK("tc_lsh_embed_kernel");
    return OOV_OK;
}

}  // namespace tc
}  // namespace oov
