// LSH OOV embedder on the tensor cores: sign-projection GEMM -> bit-pack -> multi-hot x bucket-table GEMM -> mean.
//
// Replaces inductive/torch_hash.py:55-60 (R = X P^T, bit = !(R < 0)) and inductive/lsh_embedder.py:141-179
// (out = (H W) / H.sum(1)) in ONE kernel: neither R [n, B] fp32 nor H [n, B] ever exist in HBM.
//
// Exact signs from bf16 tensor cores: every fp32 operand is split into three bf16 pieces (8 + 8 + 8 significand
// bits, an exact split), and the six products that matter are laid side by side along K:
//     A' = [x0 | x0 | x1 | x0 | x1 | x2]      B' = [p0 | p1 | p0 | p2 | p1 | p0]        (K' = 6 F, F <= 32 -> 192)
// so one K' = 192 GEMM with fp32 accumulation reproduces the fp32 projection to ~1e-7 |x||p|.  Projections closer to
// zero than 2^-17 |x| max|p| (a few per 100 000) are recomputed by the epilogue thread with the same fp32 FMA chain the
// CUDA-core path uses (csrc/lsh.cu), so both paths give identical bits; |R| < tie_eps events are counted like there.
//
// Per CTA (608 threads), persistent over 128-row tiles of the id list (tiles without OOV ids are plain row copies and
// skip the GEMMs — every role derives that from the ids with one warp vote):
//   warps 3-18  workers : fetch the NEXT tile's feature rows into registers, then per 128-plane N tile: tcgen05.ld the
//                         projections (lane = row), pack the sign bits (one bits_out word per 32-column load), count
//                         them, write the 0/1 tile H back into TENSOR MEMORY (tcgen05.st, bf16 pairs) as the A operand
//                         of the second GEMM; after the last projection split the prefetched rows into A' (128B-swizzled
//                         K-major smem) so the tensor core starts the next tile while this one is finished:
//                         out = (H W) / count (0/0 -> NaN like lsh_embedder.py:158).
//   warp 0      TMA     : B' tiles (128 planes x 64 K) through a 6-stage ring (two N tiles ahead)
//   warp 2      TMEM alloc, then TMA of the transposed bucket-table tiles (64 d x 64 planes) through an 8-stage ring
//   warp 1      MMA     : GEMM1 (SS: M128 N128 K16 x 12) into one of two TMEM accumulators; GEMM2 (TS: A = H from TMEM,
//                         M128 N64 K16 x 16) one N tile behind, accumulating H W over all N tiles in a third region.
// TMEM columns: 0-255 projections (2 buffers), 256-319 H W, 320-447 H (2 buffers).
// The fp32 bucket table is split hi + lo bf16 (two GEMM2 passes) so the sums are fp32-grade for every output dtype.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {
namespace tc {

constexpr int L_BM = 128;                 // rows per tile
constexpr int L_BN = 128;                 // planes per N tile
constexpr int L_FMAX = 32;                // features (K' = 4 F = 128 = 2 K blocks of 64)
constexpr int L_KB = 2;
constexpr int L_DMAX = 64;
constexpr int L_BSTAGES = 8, L_WSTAGES = 8;
constexpr int L_WORK_WARP0 = 3, L_WORKERS = 16;
constexpr int L_THREADS = (L_WORK_WARP0 + L_WORKERS) * 32;     // 608
constexpr int L_BT_BYTES = L_BN * 128;                         // 16 KB: one K block of one N tile of B'
constexpr int L_WT_BYTES = L_DMAX * 128;                       // 8 KB: 64 d-rows x 64 planes
constexpr int L_SMEM = 1024 + L_BSTAGES * L_BT_BYTES + L_WSTAGES * L_WT_BYTES + 6144;
constexpr float L_NEAR_REL = 3.0517578125e-5f;    // 2^-15 |x| max|p| (4 x the bound on the dropped split terms): projections
                                                  // closer to zero are recomputed in exact fp32 order

struct LshParams {
    const float* feat; int64_t n_feat_rows; int F;
    const float* planes; int B; int NT;               // NT = ceil(B / 128)
    const int64_t* ids; int64_t ids_stride; int64_t n; int64_t n_old; int64_t prime_pad;
    const void* iv_table; int iv_dtype;
    void* out; int out_dtype; int64_t out_stride; int D;
    int wsplit;                                        // 1: bf16 bucket table, 2: hi + lo
    float tie_eps;
    uint32_t* bits_out; int words;
    unsigned long long* tie_count;
    const float* pn_max;                               // largest plane norm (device scalar written by the pack kernel)
};

// ---------------------------------------------------------------- operand packing (once per call)
// Bp [NT*128, 128] bf16: row b = [p0 | p1 | p0 | p1] (32 columns each; zero for f >= F and b >= B)
// Wt [wsplit*64, NT*128] fp16: Wt[s*64 + d][b] = piece s of W[b][d] (zero padding)
__global__ void lsh_pack_kernel(const float* __restrict__ planes, int B, int F, int NT, const void* __restrict__ W, int w_dtype,
                                int D, int wsplit, __nv_bfloat16* __restrict__ Bp, __half* __restrict__ Wt,
                                float* __restrict__ pn_max) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = (int64_t)NT * L_BN;
    if (t < nb * 32) {
        const int b = (int)(t >> 5), f = (int)(t & 31);
        float p = (b < B && f < F) ? planes[(size_t)b * F + f] : 0.f;
        const __nv_bfloat16 p0 = __float2bfloat16_rn(p);
        const float r1 = p - __bfloat162float(p0);
        const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
        __nv_bfloat16* row = Bp + (size_t)b * 128;
        row[f] = p0; row[32 + f] = p1; row[64 + f] = p0; row[96 + f] = p1;
        float s2 = p * p;                                           // the 32 lanes of a warp hold one plane
        for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        if (f == 0 && s2 == s2) atomicMax(reinterpret_cast<unsigned int*>(pn_max), __float_as_uint(sqrtf(s2)));
    }
    if (t < nb * L_DMAX) {
        const int d = (int)(t / nb);
        const int64_t b = t - (int64_t)d * nb;
        const float w = (b < B && d < D) ? load_elem(W, w_dtype, b * D + d) : 0.f;
        const __half hi = __float2half_rn(w);
        Wt[(size_t)d * nb + b] = hi;
        if (wsplit == 2) Wt[(size_t)(L_DMAX + d) * nb + b] = __float2half_rn(w - __half2float(hi));
    }
}

// warp-uniform: does tile `t` (128 list positions) hold at least one OOV id?  Every role asks the same question.
__device__ __forceinline__ bool tile_has_oov(const LshParams& p, int64_t t, int lane) {
    bool any = false;
#pragma unroll
    for (int i = 0; i < L_BM / 32; ++i) {
        const int64_t r = t * L_BM + i * 32 + lane;
        if (r < p.n) any |= p.ids[r * p.ids_stride] >= p.n_old;
    }
    return __any_sync(0xffffffffu, any);
}

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(L_WORKERS * 32) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
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

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (the 0/1 tile H, bf16 pairs packed along K, lane = row) comes from
// tensor memory, where the workers put it with tcgen05.st — no shared-memory round trip for H
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_st_32x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 8 features of one row (thread = (row, slice)), fetched one tile ahead
struct Gather {
    float x[8];
    float n2;          // this slice's share of the squared row norm
};

__device__ __forceinline__ void gather_load(const LshParams& p, int64_t tile, int r, int part, Gather& gth) {
    const int64_t rr = tile * L_BM + r;
    int64_t fr = -1;
    if (rr < p.n) {
        const int64_t id = p.ids[rr * p.ids_stride];
        if (id >= p.n_old) {
            fr = feature_row(id, p.prime_pad);
            if (fr < 0 || fr >= p.n_feat_rows) fr = -1;               // out-of-range ids hash nothing (caller bug)
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int f = part * 8 + j;
        gth.x[j] = (fr >= 0 && f < p.F) ? __ldg(p.feat + fr * p.F + f) : 0.f;
        s = fmaf(gth.x[j], gth.x[j], s);
    }
    gth.n2 = s;
}

// split into two bf16 pieces and write the slice of A' = [x0 | x0 | x1 | x1] into this row's TMEM lane:
// K element k lives in column k / 2, so the 8 features are 4 columns at offset 4 * part of each 16-column segment
__device__ __forceinline__ void gather_store(uint32_t a_lane, float* sn2, int r, int part, const Gather& gth) {
    uint32_t c0[4], c1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = gth.x[2 * j], b = gth.x[2 * j + 1];
        const __nv_bfloat16 a0 = __float2bfloat16_rn(a), b0 = __float2bfloat16_rn(b);
        c0[j] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(b0) << 16);
        c1[j] = pack_bf16x2(a - __bfloat162float(a0), b - __bfloat162float(b0));
    }
    tc_st_32x4(a_lane + 0 * 16 + part * 4, c0[0], c0[1], c0[2], c0[3]);
    tc_st_32x4(a_lane + 1 * 16 + part * 4, c0[0], c0[1], c0[2], c0[3]);
    tc_st_32x4(a_lane + 2 * 16 + part * 4, c1[0], c1[1], c1[2], c1[3]);
    tc_st_32x4(a_lane + 3 * 16 + part * 4, c1[0], c1[1], c1[2], c1[3]);
    sn2[part * L_BM + r] = gth.n2;
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

__global__ void __launch_bounds__(L_THREADS, 1)
tc_lsh_embed_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmW, const LshParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sB = smem;                                        // ring of [128 planes x 64] tiles
    unsigned char* sW = sB + L_BSTAGES * L_BT_BYTES;                 // ring of [64 d x 64 planes] tiles
    unsigned char* tail = sW + L_WSTAGES * L_WT_BYTES;
    float* sn2 = reinterpret_cast<float*>(tail);                     // [4][128] squared-norm shares of the current tile's rows
    int* scnt = reinterpret_cast<int*>(sn2 + 4 * L_BM);              // [4][128] popcounts per column quarter
    uint64_t* bars = reinterpret_cast<uint64_t*>(scnt + 4 * L_BM);
    uint64_t* a_full = bars;            uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;        uint64_t* b_empty = b_full + L_BSTAGES;
    uint64_t* w_full = b_empty + L_BSTAGES;  uint64_t* w_empty = w_full + L_WSTAGES;
    uint64_t* acc1_full = w_empty + L_WSTAGES;  uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* h_full = acc1_empty + 2;  uint64_t* h_empty = h_full + 2;
    uint64_t* acc2_full = h_empty + 2;  uint64_t* acc2_empty = acc2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + L_BM - 1) / L_BM;
    const int NT = p.NT;
    const int WK = 2 * p.wsplit;                                     // Wt tiles per N tile

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmW); }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, L_WORKERS); mbar_init(a_empty, 1);
        for (int s = 0; s < L_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < L_WSTAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc1_full[a], 1); mbar_init(&acc1_empty[a], L_WORKERS);
            mbar_init(&h_full[a], L_WORKERS); mbar_init(&h_empty[a], 1);
        }
        mbar_init(acc2_full, 1); mbar_init(acc2_empty, L_WORKERS);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t ACC2_COL = 256;                               // 64 fp32 columns
    constexpr uint32_t H_COL = 320;                                  // 2 x 64 columns: [128 x 128] fp16, two per column
    constexpr uint32_t A_COL = 448;                                  // 64 columns: A' [128 x 128] bf16, two per column

    if (warp == 0) {
        // ===================== TMA: B' tiles =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0)
                for (int nt = 0; nt < NT; ++nt)
                    for (int kb = 0; kb < L_KB; ++kb) {
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&b_full[stage], L_BT_BYTES);
                        tma_load_2d(sB + stage * L_BT_BYTES, &tmB, &b_full[stage], kb * 64, nt * L_BN);
                        if (++stage == L_BSTAGES) { stage = 0; phase ^= 1; }
                    }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ===================== TMA: transposed bucket-table tiles =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0)
                for (int nt = 0; nt < NT; ++nt)
                    for (int kk = 0; kk < WK; ++kk) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&w_full[stage], L_WT_BYTES);
                        tma_load_2d(sW + stage * L_WT_BYTES, &tmW, &w_full[stage], nt * L_BN + (kk & 1) * 64, (kk >> 1) * L_DMAX);
                        if (++stage == L_WSTAGES) { stage = 0; phase ^= 1; }
                    }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = make_idesc_bf16_f32(L_BM, L_BN);
        constexpr uint32_t idesc2 = make_idesc_bf16_f32(L_BM, L_DMAX) & ~((7u << 7) | (7u << 10));   // A, B = fp16 (format 0)
        int bs = 0; uint32_t bph = 0; int ws = 0; uint32_t wph = 0;
        int64_t g1 = 0, g2 = 0;            // N tiles issued to GEMM1 / GEMM2 since kernel start
        int64_t T = 0;                     // row tiles with OOV ids done by this CTA
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0) {
                mbar_wait(a_full, (uint32_t)(T & 1));
                tc_fence_after();
                for (int nt = 0; nt <= NT; ++nt) {
                    if (nt < NT) {                                    // GEMM1(nt): projections of 128 planes
                        const int buf = (int)(g1 & 1);
                        mbar_wait(&acc1_empty[buf], (uint32_t)(((g1 >> 1) & 1) ^ 1));
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * L_BN);
                        for (int kb = 0; kb < L_KB; ++kb) {
                            mbar_wait(&b_full[bs], bph);
                            tc_fence_after();
                            const uint32_t a_tmem = tmem_base + A_COL + (uint32_t)(kb * 32);
                            const uint64_t bdesc = make_sw128_desc(smem_u32(sB + bs * L_BT_BYTES));
#pragma unroll
                            for (int k = 0; k < 4; ++k)               // A' from TMEM: K = 16 bf16 = 8 columns
                                tc_mma_bf16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), bdesc + (uint64_t)(2 * k), idesc1, (kb | k) ? 1u : 0u);
                            tc_commit(&b_empty[bs]);
                            if (++bs == L_BSTAGES) { bs = 0; bph ^= 1; }
                        }
                        tc_commit(&acc1_full[buf]);
                        if (nt == NT - 1) tc_commit(a_empty);         // A' may be rebuilt for the next row tile
                        ++g1;
                    }
                    if (nt >= 1) {                                    // GEMM2(nt - 1): acc2 += H W, H read from TMEM
                        const int j = nt - 1;
                        const int hb = (int)(g2 & 1);
                        mbar_wait(&h_full[hb], (uint32_t)((g2 >> 1) & 1));
                        if (j == 0) mbar_wait(acc2_empty, (uint32_t)((T & 1) ^ 1));
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + ACC2_COL;
                        for (int kk = 0; kk < WK; ++kk) {
                            mbar_wait(&w_full[ws], wph);
                            tc_fence_after();
                            const uint32_t a_tmem = tmem_base + H_COL + (uint32_t)(hb * 64 + (kk & 1) * 32);
                            const uint64_t bdesc = make_sw128_desc(smem_u32(sW + ws * L_WT_BYTES));
#pragma unroll
                            for (int k = 0; k < 4; ++k)               // K = 16 bf16 = 8 TMEM columns / 32 B of smem
                                tc_mma_bf16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), bdesc + (uint64_t)(2 * k), idesc2, (j | kk | k) ? 1u : 0u);
                            tc_commit(&w_empty[ws]);
                            if (++ws == L_WSTAGES) { ws = 0; wph ^= 1; }
                        }
                        tc_commit(&h_empty[hb]);
                        if (j == NT - 1) tc_commit(acc2_full);
                        ++g2;
                    }
                }
            }
            ++T;
            __syncwarp();
        }
    } else {
        // ===================== workers =====================
        const int wk = warp - L_WORK_WARP0;
        const int q = warp & 3;                    // TMEM lane quarter
        const int cq = wk >> 2;                    // column quarter of every N tile
        const int row = q * 32 + lane;             // row of the tile this thread owns in the epilogues
        const int wtid = wk * 32 + lane;           // 0..511
        const int gr = wtid >> 2, gpart = wtid & 3;   // in-vocab copy role: row, 16-column slice
        const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL;
        unsigned int my_ties = 0;
        int64_t g = 0, T = 0;
        const float pn_max = *p.pn_max;

        // first tile with OOV ids (in-vocab-only tiles on the way are plain copies)
        int64_t t = blockIdx.x;
        while (t < n_tiles && !tile_has_oov(p, t, lane)) { copy_iv_tile(p, t, gr, gpart); t += gridDim.x; }
        Gather gth;
        if (t < n_tiles) {
            gather_load(p, t, row, cq, gth);
            gather_store(a_lane, sn2, row, cq, gth);                  // a_empty: nothing has read A' yet
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        while (t < n_tiles) {
            const int64_t row0 = t * L_BM;
            worker_bar();                                             // sn2 of this tile is complete; scnt of the previous one is free
            // next tile with OOV ids: issue its gather now, it is consumed after this tile's last projection
            int64_t tn = t + gridDim.x;
            while (tn < n_tiles && !tile_has_oov(p, tn, lane)) { copy_iv_tile(p, tn, gr, gpart); tn += gridDim.x; }
            if (tn < n_tiles) gather_load(p, tn, row, cq, gth);

            int64_t my_fr = -1, my_id = INT64_MIN;
            if (row0 + row < p.n) {
                my_id = p.ids[(row0 + row) * p.ids_stride];
                if (my_id >= p.n_old) {
                    my_fr = feature_row(my_id, p.prime_pad);
                    if (my_fr < 0 || my_fr >= p.n_feat_rows) my_fr = -1;
                }
            }
            const bool my_oov = my_fr >= 0;
            // |tensor-core projection - fp32 projection| stays far below this; anything closer to zero is redone exactly
            const float xnorm = sqrtf(sn2[row] + sn2[L_BM + row] + sn2[2 * L_BM + row] + sn2[3 * L_BM + row]);
            const float near = fmaxf(L_NEAR_REL * xnorm * pn_max, 4.f * p.tie_eps);
            int cnt = 0;
            // ---- per N tile: projections -> bits -> H
            for (int nt = 0; nt < NT; ++nt, ++g) {
                const int buf = (int)(g & 1);
                const uint32_t par = (uint32_t)((g >> 1) & 1);
                mbar_wait(&acc1_full[buf], par);
                tc_fence_after();
                uint32_t v[32];
                tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * L_BN + cq * 32), v);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);
                const int b0 = nt * L_BN + cq * 32;                   // plane of column 0
                uint32_t word = 0u, nearw = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float r = __uint_as_float(v[j]);
                    word |= (r < 0.f ? 0u : 1u) << j;                 // torch_hash.py:57-59: -0, +0, NaN -> 1
                    nearw |= (fabsf(r) < near ? 1u : 0u) << j;
                }
                if (b0 + 32 > p.B) {                                  // planes past B do not exist
                    const uint32_t valid = (b0 >= p.B) ? 0u : ((1u << (p.B - b0)) - 1u);
                    word &= valid; nearw &= valid;
                }
                if (!my_oov) { word = 0u; nearw = 0u; }
                while (nearw) {                                       // rare: redo in the fp32 FMA order of csrc/lsh.cu
                    const int j = __ffs(nearw) - 1;
                    nearw &= nearw - 1;
                    const float* xr = p.feat + my_fr * p.F;
                    const float* pr = p.planes + (size_t)(b0 + j) * p.F;
                    float a = 0.f;
                    for (int f = 0; f < p.F; ++f) a = fmaf(__ldg(xr + f), __ldg(pr + f), a);
                    word = (word & ~(1u << j)) | ((a < 0.f ? 0u : 1u) << j);
                    if (fabsf(a) < p.tie_eps) ++my_ties;
                }
                cnt += __popc(word);
                if (p.bits_out != nullptr && row0 + row < p.n && nt * 4 + cq < p.words)
                    p.bits_out[(row0 + row) * p.words + nt * 4 + cq] = word;
                // H: 32 fp16 0/1 values = 16 TMEM columns of this row
                uint32_t hw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t two = (word >> (2 * i)) & 3u;
                    hw[i] = ((two & 1u) ? 0x3C00u : 0u) | ((two & 2u) ? 0x3C000000u : 0u);
                }
                mbar_wait(&h_empty[buf], par ^ 1);                    // GEMM2 of the previous use of this buffer is done
                tc_fence_after();
                tc_st_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + H_COL + (uint32_t)(buf * 64 + cq * 16), hw);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[buf]);
            }
            // ---- A' of the next tile (its features arrived long ago), so the tensor core can go on while we finish
            if (tn < n_tiles) {
                mbar_wait(a_empty, (uint32_t)(T & 1));                // this tile's GEMM1s have read A'
                tc_fence_after();
                gather_store(a_lane, sn2, row, cq, gth);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full);
            }
            // ---- final: out = (H W) / count
            scnt[cq * L_BM + row] = cnt;
            mbar_wait(acc2_full, (uint32_t)(T & 1));
            tc_fence_after();
            ++T;
            uint32_t a[16];
            tc_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + ACC2_COL + (uint32_t)(cq * 16), a);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);
            worker_bar();                                             // all four popcount partials are in smem
            const int64_t r = row0 + row;
            if (r < p.n) {
                const int d0 = cq * 16;
                const size_t osz = p.out_dtype == OOV_F32 ? 4 : 2;
                char* orow = reinterpret_cast<char*>(p.out) + (size_t)r * p.out_stride * osz;
                if (my_oov) {
                    const float den = (float)(scnt[row] + scnt[L_BM + row] + scnt[2 * L_BM + row] + scnt[3 * L_BM + row]);
                    if (p.out_dtype == OOV_BF16 && d0 + 16 <= p.D && ((reinterpret_cast<uintptr_t>(orow) + d0 * 2) & 15) == 0) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            pk[i] = pack_bf16x2(__uint_as_float(a[2 * i]) / den, __uint_as_float(a[2 * i + 1]) / den);
                        uint4* o = reinterpret_cast<uint4*>(orow + d0 * 2);
                        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (d0 + i < p.D) store_elem(p.out, p.out_dtype, r * p.out_stride + d0 + i, __uint_as_float(a[i]) / den);
                    }
                } else if (my_id != INT64_MIN && my_id >= 0 && my_id < p.n_old && p.iv_table != nullptr) {
                    for (int i = 0; i < 16; ++i)                      // in-vocab gather (bpr.py:111-112)
                        if (d0 + i < p.D)
                            store_elem(p.out, p.out_dtype, r * p.out_stride + d0 + i, load_elem(p.iv_table, p.iv_dtype, my_id * (int64_t)p.D + d0 + i));
                }
            }
            t = tn;
        }
        if (p.tie_count != nullptr) {
            for (int o = 16; o; o >>= 1) my_ties += __shfl_xor_sync(0xffffffffu, my_ties, o);
            if (lane == 0 && my_ties) atomicAdd(p.tie_count, (unsigned long long)my_ties);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------- host
bool lsh_tc_supported(int F, int B, int D) { return F >= 1 && F <= L_FMAX && D >= 1 && D <= L_DMAX && B >= 1; }

static size_t lsh_bp_bytes(int B) { return align_up((size_t)cdiv(B, L_BN) * L_BN * 128 * 2, 1024); }
static size_t lsh_wt_bytes(int B) { return align_up((size_t)2 * L_DMAX * cdiv(B, L_BN) * L_BN * 2, 1024); }
size_t lsh_tc_workspace(int B) { return lsh_bp_bytes(B) + lsh_wt_bytes(B) + 256 + 1024; }

int lsh_tc_run(const float* feat, int64_t n_feat_rows, int F, const float* planes, int B, const void* W, int w_dtype,
               const oov_rows* rows, float tie_eps, uint32_t* bits_out, unsigned long long* tie_count, void* workspace,
               size_t workspace_bytes, cudaStream_t st) {
    OOV_REQUIRE(workspace && workspace_bytes >= lsh_tc_workspace(B), OOV_ERR_WORKSPACE, "oov_lsh_embed (tcgen05): workspace %zu < %zu",
                workspace_bytes, lsh_tc_workspace(B));
    char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* Bp = reinterpret_cast<__nv_bfloat16*>(ws);
    __half* Wt = reinterpret_cast<__half*>(ws + lsh_bp_bytes(B));
    float* pn_max = reinterpret_cast<float*>(ws + lsh_bp_bytes(B) + lsh_wt_bytes(B));
    cudaError_t ce = cudaMemsetAsync(pn_max, 0, 4, st);
    OOV_REQUIRE(ce == cudaSuccess, OOV_ERR_CUDA, "cudaMemsetAsync(pn_max): %s", cudaGetErrorString(ce));
    const int NT = (int)cdiv(B, L_BN);
    const int64_t nb = (int64_t)NT * L_BN;
    LshParams p{};
    p.feat = feat; p.n_feat_rows = n_feat_rows; p.F = F; p.planes = planes; p.B = B; p.NT = NT;
    p.ids = rows->ids; p.ids_stride = rows->ids_stride; p.n = rows->n; p.n_old = rows->n_old; p.prime_pad = rows->prime_pad;
    p.iv_table = rows->iv_table; p.iv_dtype = rows->iv_dtype; p.out = rows->out; p.out_dtype = rows->out_dtype;
    p.out_stride = rows->out_stride; p.D = rows->D;
    p.wsplit = rows->out_dtype == OOV_F32 ? 2 : 1;   // fp16 hi (+ lo) pieces of the fp32 bucket table: 2^-12 (2^-23) relative
    p.tie_eps = tie_eps; p.bits_out = bits_out; p.words = (B + 31) / 32; p.tie_count = tie_count; p.pn_max = pn_max;

    const int64_t pack_threads = nb * L_DMAX > nb * 32 ? nb * L_DMAX : nb * 32;
    lsh_pack_kernel<<<(unsigned)cdiv(pack_threads, 256), 256, 0, st>>>(planes, B, F, NT, W, w_dtype, rows->D, p.wsplit, Bp, Wt, pn_max);
    OOV_LAUNCH_CHECK("lsh_pack_kernel");

    CUtensorMap tmB, tmW;
    int rc = make_tmap_bf16_2d(&tmB, Bp, 128, (uint64_t)nb, 128 * 2, L_BN);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmW, Wt, (uint64_t)nb, (uint64_t)(p.wsplit * L_DMAX), (uint64_t)nb * 2, L_DMAX);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_lsh_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM);
        OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_lsh_embed_kernel): %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int64_t n_tiles = cdiv(rows->n, L_BM);
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    tc_lsh_embed_kernel<<<grid, L_THREADS, L_SMEM, st>>>(tmB, tmW, p);
    OOV_LAUNCH_CHECK("tc_lsh_embed_kernel");
    return OOV_OK;
}

}  // namespace tc
}  // namespace oov
