// Full-sort scoring fused with masks and top-k on the tensor cores (bf16 tables, D <= 64).
//
// Replaces bpr.py:151-156 / directau.py:193-198 (score = user_e @ all_item_e.T), inductive/evaluator.py:91-94
// (pad + history -> -inf) and evaluator/collector.py:153-159 (torch.topk) in ONE kernel: the [Q, N] score matrix
// only ever exists as 128 x 256 fp32 tiles in TMEM.
//
// grid = (item CTAs, user groups of 256).  Per CTA (384 threads):
//   warp 0      TMA producer : the group's two 128-user tiles once (A, resident), then 256-item tiles (B) through
//                              a 3-stage mbarrier ring — K-major, 128B-swizzled, straight from the bf16 tables
//   warp 1      MMA issuer   : per item tile, 4 x tcgen05.mma (M 128, N 256, K 16) per user tile -> accumulator
//                              `ut` (TMEM columns ut*256 ..), then releases the smem slot
//   warp 2      TMEM alloc (512 columns)
//   warps 4-7   epilogue of user tile 0, warps 8-11 of user tile 1: thread = one user (TMEM lane); per 32-column
//               tcgen05.ld it takes a NaN-propagating 3-input max tree and compares ONCE with the user's running
//               k-th best; only chunks that beat it are scanned, masked (pad / segment / history probe) and
//               inserted into the user's sorted list (smem, thread-private column) -> epilogue cost ~1 instr/score.
// While one user tile's accumulator is being drained the tensor core fills the other one.
// Every CTA writes its per-user lists as key64 = (ordered score << 32 | ~local_row); merge_keys_kernel (topk.cu)
// reduces the per-CTA lists to the final (score desc, id asc) top-k.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {

int launch_merge_keys(const unsigned long long* partial, int P, int64_t Q, int k, int64_t off, float* out_scores,
                      int64_t* out_idx, cudaStream_t st);

namespace tc {

constexpr int SC_BM = 128;            // users per MMA (TMEM lanes)
constexpr int SC_BN = 256;            // items per tile (TMEM columns per accumulator)
constexpr int SC_UG = 256;            // users per CTA (two accumulators)
constexpr int SC_STAGES = 3;
constexpr int SC_THREADS = 384;
constexpr int SC_KMAX = 32;
constexpr int SC_A_BYTES = SC_BM * 128;               // one 128-user tile, 64 bf16 (128 B) per row
constexpr int SC_B_BYTES = SC_BN * 128;

__device__ __forceinline__ float max3_nan(float a, float b, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float max2_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ bool hist_has(const int32_t* __restrict__ cols, int lo, int hi, int64_t gid) {
    int l = lo, h = hi;
    while (l < h) {
        const int mid = (l + h) >> 1;
        if ((int64_t)cols[mid] < gid) l = mid + 1; else h = mid;
    }
    return l < hi && (int64_t)cols[l] == gid;
}

struct ScoreParams {
    int64_t Q, N;
    int k;
    int64_t item_id_offset;
    int mask_pad;
    int64_t seg_lo, seg_hi;             // global ids kept; everything else scores -inf
    const int32_t* hist_rowptr;
    const int32_t* hist_cols;
    int64_t tile_begin, tile_end;       // item tiles (of SC_BN local rows) that intersect the kept segment
    unsigned long long* partial;        // [gridDim.x][Q][k]
};

__global__ void __launch_bounds__(SC_THREADS, 1)
tc_score_topk_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmI, ScoreParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                                        // 2 x 16 KB
    unsigned char* sB = smem + 2 * SC_A_BYTES;                       // SC_STAGES x 32 KB
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(sB + SC_STAGES * SC_B_BYTES);   // [k][SC_UG]
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + (size_t)p.k * SC_UG);
    uint64_t* full_bar = bars;                  // [SC_STAGES]
    uint64_t* empty_bar = bars + SC_STAGES;     // [SC_STAGES]
    uint64_t* a_full = bars + 2 * SC_STAGES;    // [1]
    uint64_t* acc_full = a_full + 1;            // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.y * SC_UG;
    const int n_ut = (p.Q - q0 > SC_BM) ? 2 : 1;                     // valid 128-user tiles in this group

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmI); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(a_full, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(a_full, (uint32_t)(n_ut * SC_A_BYTES));
            for (int ut = 0; ut < n_ut; ++ut) tma_load_2d(sA + ut * SC_A_BYTES, &tmU, a_full, 0, (int)(q0 + ut * SC_BM));
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], SC_B_BYTES);
                tma_load_2d(sB + stage * SC_B_BYTES, &tmI, &full_bar[stage], 0, (int)(t * SC_BN));
                if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(SC_BM, SC_BN);
            mbar_wait(a_full, 0);
            int stage = 0; uint32_t phase = 0, acc_phase = 0;
            for (int64_t t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t bdesc = make_sw128_desc(smem_u32(sB + stage * SC_B_BYTES));
                for (int ut = 0; ut < n_ut; ++ut) {
                    mbar_wait(&acc_empty[ut], acc_phase ^ 1);         // this user tile's epilogue drained the accumulator
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_desc(smem_u32(sA + ut * SC_A_BYTES));
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ut * SC_BN);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
                    tc_commit(&acc_full[ut]);
                }
                tc_commit(&empty_bar[stage]);                         // item tile consumed by both user tiles
                if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int ut = (warp - 4) >> 2;                               // which user tile / accumulator
        const int q = warp & 3;                                       // TMEM lane quarter
        const int u_local = ut * SC_BM + q * 32 + lane;               // column of `lists`
        const int64_t user = q0 + u_local;
        const bool user_ok = user < p.Q;
        unsigned long long* L = lists + u_local;                      // entry e at L[e * SC_UG]
        const int k = p.k;
        for (int e = 0; e < k; ++e) L[e * SC_UG] = 0ull;
        unsigned long long thr_key = 0ull;
        float thr_f = -INFINITY;
        int hlo = 0, hhi = 0;
        if (p.hist_rowptr != nullptr && user_ok) { hlo = p.hist_rowptr[user]; hhi = p.hist_rowptr[user + 1]; }

        if (ut < n_ut) {
            uint32_t acc_phase = 0;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ut * SC_BN);
            for (int64_t t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x) {
                mbar_wait(&acc_full[ut], acc_phase);
                tc_fence_after();
                const int64_t row0 = t * SC_BN;                       // first local item row of the tile
#pragma unroll 1
                for (int c = 0; c < SC_BN / 32; ++c) {
                    uint32_t v[32];
                    tc_ld_32x32(t_lane + (uint32_t)(c * 32), v);
                    tc_wait_ld();
                    // NaN-propagating max of the 32 scores: 11 + 4 + 1 three-input / two-input max ops
                    float m[11];
#pragma unroll
                    for (int j = 0; j < 10; ++j)
                        m[j] = max3_nan(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
                    m[10] = max2_nan(__uint_as_float(v[30]), __uint_as_float(v[31]));
                    const float m0 = max3_nan(m[0], m[1], m[2]), m1 = max3_nan(m[3], m[4], m[5]);
                    const float m2 = max3_nan(m[6], m[7], m[8]), m3 = max2_nan(m[9], m[10]);
                    const float mx = max2_nan(max3_nan(m0, m1, m2), m3);
                    if (!(mx < thr_f) && user_ok) {                   // rare once the list has warmed up
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float s = __uint_as_float(v[j]);
                            if (s < thr_f) continue;
                            const int64_t li = row0 + c * 32 + j;
                            if (li >= p.N) continue;                  // zero-filled rows past the end of the shard
                            const int64_t gid = li + p.item_id_offset;
                            if ((p.mask_pad && gid == 0) || gid < p.seg_lo || gid >= p.seg_hi ||
                                (hhi > hlo && hist_has(p.hist_cols, hlo, hhi, gid)))
                                s = -INFINITY;
                            const unsigned long long key = make_key64(s, (uint32_t)li);
                            if (key > thr_key) {
                                int e = k - 1;
                                while (e > 0) {
                                    const unsigned long long prev = L[(e - 1) * SC_UG];
                                    if (prev >= key) break;
                                    L[e * SC_UG] = prev;
                                    --e;
                                }
                                L[e * SC_UG] = key;
                                thr_key = L[(k - 1) * SC_UG];
                                thr_f = thr_key ? key64_score(thr_key) : -INFINITY;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ut]);
                acc_phase ^= 1;
            }
        }
        if (user_ok) {
            unsigned long long* dst = p.partial + ((size_t)blockIdx.x * p.Q + user) * k;
            for (int e = 0; e < k; ++e) dst[e] = L[e * SC_UG];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static size_t score_smem_bytes(int k) {
    return 1024 + 2 * SC_A_BYTES + SC_STAGES * SC_B_BYTES + (size_t)k * SC_UG * 8 + 256;
}

bool score_tc_supported(int dtype, int D, int k) { return dtype == OOV_BF16 && D >= 8 && D <= 64 && D % 8 == 0 && k >= 1 && k <= SC_KMAX; }

static int score_grid_x(int64_t Q, int64_t n_tiles) {
    const int64_t groups = cdiv(Q, SC_UG);
    int64_t gx = num_sms() / groups;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    return (int)gx;
}

size_t score_tc_workspace(int64_t Q, int64_t N, int k) {
    const int gx = score_grid_x(Q, cdiv(N > 0 ? N : 1, SC_BN));
    return align_up((size_t)gx * Q * k * 8, 256);
}

int score_tc_run(const void* users, const void* items, int64_t Q, int64_t N, int D, int k, int64_t item_id_offset,
                 int mask_pad, int64_t seg_lo, int64_t seg_hi, const int32_t* hist_rowptr, const int32_t* hist_cols,
                 float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    // only item tiles that intersect the kept segment are visited
    int64_t lo = seg_lo - item_id_offset, hi = seg_hi > item_id_offset + N ? N : seg_hi - item_id_offset;
    if (lo < 0) lo = 0;
    if (hi > N) hi = N;
    ScoreParams p{};
    p.Q = Q; p.N = N; p.k = k; p.item_id_offset = item_id_offset; p.mask_pad = mask_pad;
    p.seg_lo = seg_lo; p.seg_hi = seg_hi; p.hist_rowptr = hist_rowptr; p.hist_cols = hist_cols;
    p.tile_begin = hi > lo ? lo / SC_BN : 0;
    p.tile_end = hi > lo ? cdiv(hi, SC_BN) : 0;
    const int64_t n_tiles = p.tile_end - p.tile_begin;
    const int gx = score_grid_x(Q, n_tiles > 0 ? n_tiles : 1);
    const size_t need = (size_t)gx * Q * k * 8;
    OOV_REQUIRE(workspace && workspace_bytes >= need, OOV_ERR_WORKSPACE, "oov_fullsort_topk (tcgen05): workspace %zu < %zu",
                workspace_bytes, need);
    p.partial = reinterpret_cast<unsigned long long*>(workspace);

    CUtensorMap tmU, tmI;
    int rc = make_tmap_bf16_2d(&tmU, users, (uint64_t)D, (uint64_t)Q, (uint64_t)D * 2, SC_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmI, items, (uint64_t)D, (uint64_t)(N > 0 ? N : 1), (uint64_t)D * 2, SC_BN);
    if (rc) return rc;
    const size_t smem = score_smem_bytes(k);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)score_smem_bytes(SC_KMAX));
        OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_score_topk_kernel): %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const dim3 grid((unsigned)gx, (unsigned)cdiv(Q, SC_UG));
    tc_score_topk_kernel<<<grid, SC_THREADS, smem, st>>>(tmU, tmI, p);
    OOV_LAUNCH_CHECK("tc_score_topk_kernel");
    return launch_merge_keys(p.partial, gx, Q, k, item_id_offset, out_scores, out_idx, st);
}

}  // namespace tc
}  // namespace oov
