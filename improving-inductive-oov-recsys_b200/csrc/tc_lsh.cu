// LSH OOV embedder on the tensor cores: sign-projection GEMM -> signs -> multi-hot x bucket-table GEMM -> mean.
//
// Replaces inductive/torch_hash.py:55-60 (R = X P^T, bit = !(R < 0)) and inductive/lsh_embedder.py:141-179
// (out = (H W) / H.sum(1)) in ONE kernel: neither R [n, B] fp32 nor H [n, B] ever exist in HBM or shared memory.
//
// Exact signs from bf16 tensor cores: x = x0 + x1 + (<= 2^-18 |x|), p likewise (bf16 pieces), and the three products
// that matter sit side by side along K:
//     A' = [x0 | x0 | x1]      B' = [p0 | p1 | p0]        (K' = 3 F, F <= 32 -> 96)
// The dropped terms are bounded by 1.5 * 2^-17 |x||p|; projections closer to zero than 2^-16 |x| max|p| (about 1 per
// 10 000) are recomputed by the epilogue thread with the fp32 FMA chain of the CUDA-core path (csrc/lsh.cu), so both
// paths give identical bits, and |R| < tie_eps events are counted like there.
//
// Second GEMM without popcounts or masks: the epilogue writes S' = 2H - 1 (+-1 in fp16: the sign bit of R under a
// constant) and the bucket table gets one extra column of ones (zero for planes >= B), so
//     S' [W | 1] = [2 H W - colsum(W) | 2 count - B]   ->   out = (acc + colsum(W)) / (acc_ones + B)
// with 0 / 0 -> NaN like lsh_embedder.py:158.  W is split into fp16 hi (+ lo for fp32 outputs) pieces.
//
// Per CTA (640 threads), persistent over 128-row tiles of the id list (tiles without OOV ids are plain row copies and
// skip the GEMMs — every role derives that from the ids with one warp vote):
//   warps 4-19  workers : fetch the NEXT tile's feature rows into registers; per 128-plane N tile: tcgen05.ld the
//                         projections (lane = row), min|R| tree against the near-zero threshold, two instructions
//                         per pair of scores (PRMT + LOP3) to form the fp16 +-1 words, tcgen05.st them back into
//                         TENSOR MEMORY as the A operand of the second GEMM; after the last projection split the
//                         prefetched rows into A' (also in TMEM) so the tensor core starts the next tile at once.
//   warp 0      TMA     : B' tiles (128 planes x 128 K) through a 4-stage ring (one mbarrier per N tile)
//   warp 2      TMEM alloc, then TMA of the transposed bucket-table tiles (80 rows x 128 planes) through a 3-stage ring
//   warp 1      MMA 1   : GEMM1 (TS: A' from TMEM, M128 N128 K16 x 6) into one of two TMEM accumulators
//   warp 3      MMA 2   : GEMM2 (TS: A = S' from TMEM, M128 N80 K16 x 8 per piece), accumulating over all N tiles.
//                         Two issuing threads: each spends ~100 cycles per mbarrier wait, one thread could not keep up.
// TMEM columns: 0-255 projections (2 buffers), 256-335 [S' W | S' 1], 336-463 S' (2 buffers), 464-511 A'.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {
namespace tc {

constexpr int L_BM = 128;                 // rows per tile
constexpr int L_BN = 128;                 // planes per N tile
constexpr int L_FMAX = 32;                // features (K' = 3 F = 96: K block 0 full, K block 1 half used)
constexpr int L_DMAX = 64;
constexpr int L_WROWS = 80;               // bucket-table tile rows: 64 d + the ones row + padding to a multiple of 16
constexpr int L_BSTAGES = 4, L_WSTAGES = 3;   // one stage = one N tile of B' (2 K blocks) / one piece of the bucket tile (2 halves)
constexpr int L_WORK_WARP0 = 4, L_WORKERS = 16;
constexpr int L_THREADS = (L_WORK_WARP0 + L_WORKERS) * 32;     // 640
constexpr int L_BT_BYTES = 2 * L_BN * 128;                     // 32 KB: both K blocks of one N tile of B'
constexpr int L_WT_BYTES = 2 * L_WROWS * 128;                  // 20 KB: 80 rows x 128 planes (two 64-plane halves)
constexpr int L_SMEM = 1024 + L_BSTAGES * L_BT_BYTES + L_WSTAGES * L_WT_BYTES + 4096;
constexpr float L_NEAR_REL = 1.52587890625e-5f;   // 2^-16 |x| max|p|: projections closer to zero are recomputed in fp32 order

struct LshParams {
    const float* feat; int64_t n_feat_rows; int F;
    const float* planes; int B; int NT;               // NT = ceil(B / 128)
    const int64_t* ids; int64_t ids_stride; int64_t n; int64_t n_old; int64_t prime_pad;
    const void* iv_table; int iv_dtype;
    void* out; int out_dtype; int64_t out_stride; int D;
    int wsplit;                                        // 1: fp16 bucket table, 2: hi + lo
    float tie_eps;
    uint32_t* bits_out; int words;
    unsigned long long* tie_count;
    const float* pn_max;                               // largest plane norm (device scalar written by the pack kernel)
    const float* wsum;                                 // [64] column sums of the packed bucket table
};

// ---------------------------------------------------------------- operand packing (once per call)
// Bp [NT*128, 128] bf16: row b = [p0 | p1 | p0 | 0] (32 columns each; zero for f >= F and b >= B)
// Wt [wsplit*80, NT*128] fp16: rows s*80 + d = piece s of W[., d]; row 64 of piece 0 = 1 for b < B; zero padding
__global__ void lsh_pack_kernel(const float* __restrict__ planes, int B, int F, int NT, const void* __restrict__ W, int w_dtype,
                                int D, int wsplit, __nv_bfloat16* __restrict__ Bp, __half* __restrict__ Wt,
                                float* __restrict__ pn_max) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = (int64_t)NT * L_BN;
    if (t < nb * 32) {
        const int b = (int)(t >> 5), f = (int)(t & 31);
        float p = (b < B && f < F) ? planes[(size_t)b * F + f] : 0.f;
        const __nv_bfloat16 p0 = __float2bfloat16_rn(p);
        const __nv_bfloat16 p1 = __float2bfloat16_rn(p - __bfloat162float(p0));
        __nv_bfloat16* row = Bp + (size_t)b * 128;
        row[f] = p0; row[32 + f] = p1; row[64 + f] = p0; row[96 + f] = __float2bfloat16_rn(0.f);
        float s2 = p * p;                                           // the 32 lanes of a warp hold one plane
        for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        if (f == 0 && s2 == s2) atomicMax(reinterpret_cast<unsigned int*>(pn_max), __float_as_uint(sqrtf(s2)));
    }
    if (t < nb * L_WROWS) {
        const int d = (int)(t / nb);
        const int64_t b = t - (int64_t)d * nb;
        float w = 0.f;
        if (b < B) w = d < D ? load_elem(W, w_dtype, b * D + d) : (d == L_DMAX ? 1.f : 0.f);
        const __half hi = __float2half_rn(w);
        Wt[(size_t)d * nb + b] = hi;
        if (wsplit == 2) Wt[(size_t)(L_WROWS + d) * nb + b] = __float2half_rn(w - __half2float(hi));
    }
}
// wsum[d] = sum over b of the packed pieces of W[b, d], fixed order (one warp per column, fp32 tree over 32 partial sums)
__global__ void lsh_wsum_kernel(const __half* __restrict__ Wt, int64_t nb, int wsplit, float* __restrict__ wsum) {
    const int d = blockIdx.x, lane = threadIdx.x;
    float s = 0.f;
    for (int64_t b = lane; b < nb; b += 32) {
        float w = __half2float(Wt[(size_t)d * nb + b]);
        if (wsplit == 2) w += __half2float(Wt[(size_t)(L_WROWS + d) * nb + b]);
        s += w;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[d] = s;
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
__device__ __forceinline__ uint32_t tc_ld_32x1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
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

// split into two bf16 pieces and write the slice of A' = [x0 | x0 | x1] into this row's TMEM lane:
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
    unsigned char* sW = sB + L_BSTAGES * L_BT_BYTES;                 // ring of [80 rows x 64 planes] tiles
    unsigned char* tail = sW + L_WSTAGES * L_WT_BYTES;
    float* sn2 = reinterpret_cast<float*>(tail);                     // [4][128] squared-norm shares of the current tile's rows
    float* swsum = sn2 + 4 * L_BM;                                   // [64] column sums of the bucket table
    uint64_t* bars = reinterpret_cast<uint64_t*>(swsum + L_DMAX);
    uint64_t* a_full = bars;            uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;        uint64_t* b_empty = b_full + L_BSTAGES;
    uint64_t* w_full = b_empty + L_BSTAGES;  uint64_t* w_empty = w_full + L_WSTAGES;
    uint64_t* acc1_full = w_empty + L_WSTAGES;  uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* h_full = acc1_empty + 2;  uint64_t* h_empty = h_full + 2;
    uint64_t* acc2_full = h_empty + 2;  uint64_t* acc2_empty = acc2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

    // Logical warp ids are the physical ones rotated by L_WORK_WARP0 (a multiple of 4, so `warp & 3` is still the TMEM lane
    // quarter): the single-thread TMA / MMA roles (logical 0-3) run on the HIGHEST physical warps, which the
    // sub-partition schedulers favour over the sixteen epilogue warps (B300_MICROARCH.md: highest warp id first).
    const int warp = (int)((threadIdx.x >> 5) + L_WORK_WARP0) % (int)(L_THREADS / 32), lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + L_BM - 1) / L_BM;
    const int NT = p.NT;

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
    if (threadIdx.x >= 128 && threadIdx.x < 128 + L_DMAX) swsum[threadIdx.x - 128] = p.wsum[threadIdx.x - 128];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t ACC2_COL = 256;                               // 80 fp32 columns: [S' W | S' 1 | padding]
    constexpr uint32_t H_COL = 336;                                  // 2 x 64 columns: S' [128 x 128] fp16, two per column
    constexpr uint32_t A_COL = 464;                                  // 48 columns: A' [128 x 96] bf16, two per column

    if (warp == 0) {
        // ===================== TMA: B' tiles =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0)
                for (int nt = 0; nt < NT; ++nt) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&b_full[stage], L_BT_BYTES);
                    tma_load_2d(sB + stage * L_BT_BYTES, &tmB, &b_full[stage], 0, nt * L_BN);
                    tma_load_2d(sB + stage * L_BT_BYTES + L_BT_BYTES / 2, &tmB, &b_full[stage], 64, nt * L_BN);
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
                    for (int pc = 0; pc < p.wsplit; ++pc) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&w_full[stage], L_WT_BYTES);
                        tma_load_2d(sW + stage * L_WT_BYTES, &tmW, &w_full[stage], nt * L_BN, pc * L_WROWS);
                        tma_load_2d(sW + stage * L_WT_BYTES + L_WT_BYTES / 2, &tmW, &w_full[stage], nt * L_BN + 64, pc * L_WROWS);
                        if (++stage == L_WSTAGES) { stage = 0; phase ^= 1; }
                    }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer 1: projections =====================
        constexpr uint32_t idesc1 = make_idesc_bf16_f32(L_BM, L_BN);
        int bs = 0; uint32_t bph = 0;
        int64_t g1 = 0;                    // N tiles issued since kernel start
        int64_t T = 0;                     // row tiles with OOV ids done by this CTA
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0) {
                mbar_wait(a_full, (uint32_t)(T & 1));
                for (int nt = 0; nt < NT; ++nt, ++g1) {
                    const int buf = (int)(g1 & 1);
                    mbar_wait(&acc1_empty[buf], (uint32_t)(((g1 >> 1) & 1) ^ 1));
                    mbar_wait(&b_full[bs], bph);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * L_BN);
                    const uint32_t a_tmem = tmem_base + A_COL;
                    const uint64_t bdesc0 = make_sw128_desc(smem_u32(sB + bs * L_BT_BYTES));
                    const uint64_t bdesc1 = make_sw128_desc(smem_u32(sB + bs * L_BT_BYTES + L_BT_BYTES / 2));
#pragma unroll
                    for (int k = 0; k < 6; ++k)                       // A' from TMEM: K = 16 bf16 = 8 columns; K' = 96 -> 6 steps
                        tc_mma_f16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), (k < 4 ? bdesc0 + (uint64_t)(2 * k) : bdesc1 + (uint64_t)(2 * (k - 4))),
                                      idesc1, k ? 1u : 0u);
                    tc_commit(&b_empty[bs]);
                    tc_commit(&acc1_full[buf]);
                    if (nt == NT - 1) tc_commit(a_empty);             // A' may be rebuilt for the next row tile
                    if (++bs == L_BSTAGES) { bs = 0; bph ^= 1; }
                }
            }
            ++T;
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===================== MMA issuer 2: acc2 += S' [W | 1], S' read from TMEM =====================
        constexpr uint32_t idesc2 = make_idesc_bf16_f32(L_BM, L_WROWS) & ~((7u << 7) | (7u << 10));   // A, B = fp16 (format 0)
        int ws = 0; uint32_t wph = 0;
        int64_t g2 = 0;
        int64_t T = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (!tile_has_oov(p, t, lane)) continue;
            if (lane == 0) {
                const uint32_t d_tmem = tmem_base + ACC2_COL;
                for (int j = 0; j < NT; ++j, ++g2) {
                    const int hb = (int)(g2 & 1);
                    mbar_wait(&h_full[hb], (uint32_t)((g2 >> 1) & 1));
                    if (j == 0) mbar_wait(acc2_empty, (uint32_t)((T & 1) ^ 1));
                    for (int pc = 0; pc < p.wsplit; ++pc) {
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        const uint32_t a_tmem = tmem_base + H_COL + (uint32_t)(hb * 64);
                        const uint64_t bdesc0 = make_sw128_desc(smem_u32(sW + ws * L_WT_BYTES));
                        const uint64_t bdesc1 = make_sw128_desc(smem_u32(sW + ws * L_WT_BYTES + L_WT_BYTES / 2));
#pragma unroll
                        for (int k = 0; k < 8; ++k)                   // K = 16 fp16 = 8 TMEM columns / 32 B of smem
                            tc_mma_f16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), (k < 4 ? bdesc0 + (uint64_t)(2 * k) : bdesc1 + (uint64_t)(2 * (k - 4))),
                                          idesc2, (j | pc | k) ? 1u : 0u);
                        tc_commit(&w_empty[ws]);
                        if (++ws == L_WSTAGES) { ws = 0; wph ^= 1; }
                    }
                    tc_commit(&h_empty[hb]);
                    if (j == NT - 1) tc_commit(acc2_full);
                }
            }
            ++T;
            __syncwarp();
        }
    } else {
        // ===================== workers =====================
        const int wk = warp - L_WORK_WARP0;
        const int q = warp & 3;                    // TMEM lane quarter
        const int cq = wk >> 2;                    // column quarter of every N tile / 8-feature slice / 16-column output slice
        const int row = q * 32 + lane;             // row of the tile this thread owns
        const int wtid = wk * 32 + lane;           // 0..511
        const int gr = wtid >> 2, gpart = wtid & 3;   // in-vocab copy role: row, 16-column slice
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a_lane = lane_base + A_COL;
        unsigned int my_ties = 0;
        int64_t g = 0, T = 0;
        const float pn_max = *p.pn_max;
        const uint32_t SIGNS = 0x80008000u, ONES = 0x3C003C00u;       // fp16 pair: sign bits / (+1, +1)

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
            worker_bar();                                             // sn2 of this tile is complete
            // next tile with OOV ids: issue its gather now, it is consumed after this tile's last projection
            int64_t tn = t + gridDim.x;
            while (tn < n_tiles && !tile_has_oov(p, tn, lane)) { copy_iv_tile(p, tn, gr, gpart); tn += gridDim.x; }
            const float xnorm = sqrtf(sn2[row] + sn2[L_BM + row] + sn2[2 * L_BM + row] + sn2[3 * L_BM + row]);
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
            const float near = fmaxf(L_NEAR_REL * xnorm * pn_max, 4.f * p.tie_eps);
            // ---- per N tile: projections -> signs -> S'
            for (int nt = 0; nt < NT; ++nt, ++g) {
                const int buf = (int)(g & 1);
                const uint32_t par = (uint32_t)((g >> 1) & 1);
                mbar_wait(&acc1_full[buf], par);
                tc_fence_after();
                uint32_t v[32];
                tc_ld_32x32(lane_base + (uint32_t)(buf * L_BN + cq * 32), v);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);
                // fast path: S' pair = (+1, +1) with the sign bits of the two projections (2 instructions per pair)
                uint32_t hw[16];
                float mn = INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    hw[i] = (__byte_perm(v[2 * i], v[2 * i + 1], 0x7030) & SIGNS) ^ ONES;
                    mn = fminf(mn, fminf(fabsf(__uint_as_float(v[2 * i])), fabsf(__uint_as_float(v[2 * i + 1]))));
                }
                const int b0 = nt * L_BN + cq * 32;                   // plane of column 0
                if ((mn < near || p.bits_out != nullptr) && my_oov) {
                    // slow path (about 1 chunk in 200, or when the caller wants the multi-hot words): exact bits.
                    // -0 has the sign bit but is not < 0 (torch_hash.py:57-59), near-zero projections are redone in
                    // the fp32 FMA order of csrc/lsh.cu
                    uint32_t word = 0u, nearw = 0u;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float r = __uint_as_float(v[j]);
                        word |= (r < 0.f ? 0u : 1u) << j;
                        nearw |= (fabsf(r) < near ? 1u : 0u) << j;
                    }
                    const uint32_t valid = (b0 + 32 <= p.B) ? 0xffffffffu : ((b0 >= p.B) ? 0u : ((1u << (p.B - b0)) - 1u));
                    nearw &= valid;
                    while (nearw) {
                        const int j = __ffs(nearw) - 1;
                        nearw &= nearw - 1;
                        const float* xr = p.feat + my_fr * p.F;
                        const float* pr = p.planes + (size_t)(b0 + j) * p.F;
                        float a = 0.f;
                        for (int f = 0; f < p.F; ++f) a = fmaf(__ldg(xr + f), __ldg(pr + f), a);
                        word = (word & ~(1u << j)) | ((a < 0.f ? 0u : 1u) << j);
                        if (fabsf(a) < p.tie_eps) ++my_ties;
                    }
                    word &= valid;
                    if (p.bits_out != nullptr && row0 + row < p.n && nt * 4 + cq < p.words)
                        p.bits_out[(row0 + row) * p.words + nt * 4 + cq] = word;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {                    // planes >= B meet zero bucket rows: any sign will do
                        const uint32_t two = (word >> (2 * i)) & 3u;
                        hw[i] = ((two & 1u) ? 0x3C00u : 0xBC00u) | ((two & 2u) ? 0x3C000000u : 0xBC000000u);
                    }
                } else if (p.bits_out != nullptr && row0 + row < p.n && nt * 4 + cq < p.words) {
                    p.bits_out[(row0 + row) * p.words + nt * 4 + cq] = 0u;
                }
                mbar_wait(&h_empty[buf], par ^ 1);                    // GEMM2 of the previous use of this buffer is done
                tc_fence_after();
                tc_st_32x16(lane_base + H_COL + (uint32_t)(buf * 64 + cq * 16), hw);
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
            // ---- final: out = (S' W + colsum W) / (S' 1 + B)   [= 2 H W / 2 count]
            mbar_wait(acc2_full, (uint32_t)(T & 1));
            tc_fence_after();
            ++T;
            uint32_t a[16];
            tc_ld_32x16(lane_base + ACC2_COL + (uint32_t)(cq * 16), a);
            const uint32_t ones_acc = tc_ld_32x1(lane_base + ACC2_COL + L_DMAX);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);
            const int64_t r = row0 + row;
            if (r < p.n) {
                const int d0 = cq * 16;
                const size_t osz = p.out_dtype == OOV_F32 ? 4 : 2;
                char* orow = reinterpret_cast<char*>(p.out) + (size_t)r * p.out_stride * osz;
                if (my_oov) {
                    const float den = __uint_as_float(ones_acc) + (float)p.B;          // 2 x count, exact
                    float o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        o[i] = den == 0.f ? __uint_as_float(0x7FC00000u) : (__uint_as_float(a[i]) + swsum[d0 + i]) / den;
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
static size_t lsh_wt_bytes(int B) { return align_up((size_t)2 * L_WROWS * cdiv(B, L_BN) * L_BN * 2, 1024); }
size_t lsh_tc_workspace(int B) { return lsh_bp_bytes(B) + lsh_wt_bytes(B) + 512 + 1024; }

int lsh_tc_run(const float* feat, int64_t n_feat_rows, int F, const float* planes, int B, const void* W, int w_dtype,
               const oov_rows* rows, float tie_eps, uint32_t* bits_out, unsigned long long* tie_count, void* workspace,
               size_t workspace_bytes, cudaStream_t st) {
    OOV_REQUIRE(workspace && workspace_bytes >= lsh_tc_workspace(B), OOV_ERR_WORKSPACE, "oov_lsh_embed (tcgen05): workspace %zu < %zu",
                workspace_bytes, lsh_tc_workspace(B));
    char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* Bp = reinterpret_cast<__nv_bfloat16*>(ws);
    __half* Wt = reinterpret_cast<__half*>(ws + lsh_bp_bytes(B));
    float* pn_max = reinterpret_cast<float*>(ws + lsh_bp_bytes(B) + lsh_wt_bytes(B));
    float* wsum = pn_max + 16;
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
    p.tie_eps = tie_eps; p.bits_out = bits_out; p.words = (B + 31) / 32; p.tie_count = tie_count; p.pn_max = pn_max; p.wsum = wsum;

    lsh_pack_kernel<<<(unsigned)cdiv(nb * L_WROWS, 256), 256, 0, st>>>(planes, B, F, NT, W, w_dtype, rows->D, p.wsplit, Bp, Wt, pn_max);
    OOV_LAUNCH_CHECK("lsh_pack_kernel");
    lsh_wsum_kernel<<<L_DMAX, 32, 0, st>>>(Wt, nb, p.wsplit, wsum);
    OOV_LAUNCH_CHECK("lsh_wsum_kernel");

    CUtensorMap tmB, tmW;
    int rc = make_tmap_bf16_2d(&tmB, Bp, 128, (uint64_t)nb, 128 * 2, L_BN);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmW, Wt, (uint64_t)nb, (uint64_t)(p.wsplit * L_WROWS), (uint64_t)nb * 2, L_WROWS);   // fp16: same 2-byte boxes
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
