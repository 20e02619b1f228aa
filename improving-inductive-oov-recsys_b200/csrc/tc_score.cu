// Full-sort scoring fused with masks and top-k on the tensor cores (bf16 tables, D <= 64).
//
// Replaces bpr.py:151-156 / directau.py:193-198 (score = user_e @ all_item_e.T), inductive/evaluator.py:91-94
// (pad + history -> -inf) and evaluator/collector.py:153-159 (torch.topk) in ONE kernel: the [Q, N] score matrix
// only ever exists as 128 x 128 fp32 tiles in TMEM.
//
// grid = (item ranges, user groups of 512).  Every CTA owns a CONTIGUOUS range of 128-item tiles.  608 threads:
//   warp 16      TMA producer : the group's four 128-user tiles once (A, resident), then 128-item tiles (B) through
//                               an mbarrier ring — K-major, 128B-swizzled, straight from the bf16 tables
//   warps 17-18  MMA issuers  : two user tiles `ut` each; per item tile and user tile 4 x tcgen05.mma (M 128, N 128, K 16) into
//                               accumulator `ut` (TMEM columns ut*128 ..); the tensor core works on the other
//                               three user tiles while one is being drained
//   warps 0-15   epilogue     : thread = one user (TMEM lane).  Per 32-column tcgen05.ld a NaN-propagating 3-input max
//                               tree is compared ONCE with the user's threshold (one warp vote, a second vote per
//                               quarter chunk); only what beats it is queued, masked (pad / segment / history) and
//                               inserted into the user's list (smem, thread-private column).
// Threshold sharing (shards below 256 full tiles; larger shards take the sampled pre-pass described below): a user is
// scored by P = gridDim.x CTAs ("streams"), and a short stream's own k-th best is a weak filter.  Every stream publishes its j-th best score (j = ceil(k / G), G = min(P, k)) with a plain store;
// the streams are dealt into G groups and T = min over groups of (max over the group's streams of the published
// value) has at least G * j >= k distinct scores at or above it, so dropping scores strictly below T is exact.
// Threads refresh T on a doubling schedule (after tiles 1, 2, 4, 8, ...): total candidates per user stay near
// k' ln(N) for the whole grid instead of P k ln(N / P).
// History: items are visited in increasing id order, so each thread walks its user's sorted CSR history with a
// cursor (one register holds the next masked id) instead of searching.
// Every CTA writes its per-user lists as key64 = (ordered score << 32 | ~local_row); merge_keys_kernel (topk.cu)
// reduces the per-CTA lists to the final (score desc, id asc) top-k.
//
// Sampled pre-pass (large shards): the same TMA / tcgen05 pipeline first visits every `stride`-th FULL item tile with a
// divergence-free epilogue that only records each user's maximum score of the tile (MODE 1).  score_threshold_kernel
// then takes, per user, the R-th largest tile maximum with R = k + (masked items of that user inside sampled tiles):
// every masked item is the maximum of at most one tile, so at least k sampled tiles have an UNMASKED item at or above
// that value and no score strictly below it can be in the top-k.  The main pass (MODE 0) starts from this threshold
// (pass rate ~ R / sampled items, e.g. 1e-4 for 1 M items at stride 4) instead of -inf: the fill phase, most list
// maintenance and the cross-CTA threshold exchange disappear, for `1 / stride` extra MMA work.  Both passes issue the
// identical MMA sequence for a (user tile, item tile) pair, so a score has the same bits in both.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {

int launch_merge_keys(const unsigned long long* partial, const uint8_t* partial_n, int P, int64_t Q, int k, int64_t off,
                      float* out_scores, int64_t* out_idx, cudaStream_t st, const KeyOut ko);

namespace tc {

constexpr int SC_BM = 128;            // users per MMA (TMEM lanes)
constexpr int SC_BN = 128;            // items per tile (TMEM columns per accumulator)
constexpr int SC_NUT = 4;             // user tiles (accumulators) per CTA
constexpr int SC_UG = SC_BM * SC_NUT; // users per CTA
constexpr int SC_MAX_STAGES = 10;
constexpr int SC_L2_AHEAD = 12;          // item tiles the producer asks L2 to fetch ahead of the TMA ring
// Warp roles.  The SM sub-partition schedulers favour the highest warp id among eligible warps, so the two latency-
// critical single-thread roles (TMA producer, MMA issuer) take the HIGHEST ids: as warps 0 / 1 they were starved by
// the sixteen epilogue warps whenever those had work or polled a barrier.
constexpr int SC_EPI_WARP0 = 0;
constexpr int SC_EPI_WARPS = 4 * SC_NUT;
constexpr int SC_W_TMA = SC_EPI_WARPS;                         // 16: TMEM alloc / dealloc + TMA producer
constexpr int SC_W_ALLOC = SC_W_TMA;
constexpr int SC_W_MMA = SC_EPI_WARPS + 1;                     // 17, 18: MMA issuers, two user tiles each
constexpr int SC_THREADS = (SC_EPI_WARPS + 3) * 32;            // 608
constexpr int SC_KMAX = 24;              // smem: lists k x 512 x 8 B next to A (64 KB), the queues (32 KB) and >= 2 B stages
constexpr int SC_PUB_MAXP = 160;          // streams per user the threshold refresh reads (>= SM count)
constexpr int SC_QCAP = 8;               // queued candidates per user between drains
constexpr int SC_A_BYTES = SC_BM * 128;               // one 128-user tile, 64 bf16 (128 B) per row
constexpr int SC_B_BYTES = SC_BN * 128;
constexpr int SC_SMEM_MAX = 232448;                   // 227 KB opt-in limit

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

struct ScoreParams {
    int64_t Q, N;
    int k, stages;
    int64_t item_id_offset;
    int mask_pad;
    int64_t seg_lo, seg_hi;             // global ids kept; everything else scores -inf
    const int32_t* hist_rowptr;
    const int32_t* hist_cols;           // sorted per user
    int64_t tile_begin, tile_end;       // item tiles (of SC_BN local rows) that intersect the kept segment
    uint32_t* pub;                      // [Q][gridDim.x] j-th best score of each stream (ordered-float key; 0 = none yet)
    int pub_j, pub_groups;              // j (1..3, or k when P is too small) and G of the threshold-sharing scheme
    int debug;                          // profiling only (OOV_SCORE_DEBUG): bit 0 = no candidate processing, bit 1 = no threshold sharing
    unsigned long long* partial;        // [gridDim.x][Q][k]: the first partial_n[stream][user] entries are valid
    uint8_t* partial_n;                 // [gridDim.x][Q]
    int64_t tile_stride;                // MODE 1: tiles tile_begin + i * tile_stride, i in [0, n_visit)
    int64_t n_visit;                    // tiles visited by the whole grid (MODE 0: tile_end - tile_begin)
    uint32_t* tile_max;                 // MODE 1 out: [Q][n_visit] ordered-float key of the user's best score in tile i
    const uint32_t* thr_init;           // MODE 0 in (optional): [Q] key no top-k score is below (0 = none)
    int share;                          // MODE 0: exchange thresholds between the CTAs of a user
    const uint32_t* choice;             // optional device flag written by score_choose_kernel: 1 = the column-split main pass runs,
                                        // 0 = the thread-per-user one; the kernel that is not chosen returns at once
};

// per-thread (= per-user) epilogue state.  The list is UNSORTED and split in two u32 arrays (ordered score / ~row):
// entries [0, n) are valid.  While n < k new candidates are appended; when it is full the smallest entry (tracked in
// thr_hi / thr_lo / min_pos) is replaced and the minimum is searched again with short independent chains.
struct UserState {
    uint32_t* Lh;                       // ordered score keys: entry e at Lh[e * SC_UG]
    uint32_t* Ll;                       // ~local row:         entry e at Ll[e * SC_UG]
    unsigned long long* Qe;             // candidate queue column: entry i at Qe[i * SC_UG] = (score bits << 32 | local row)
    uint32_t thr_hi, thr_lo;            // smallest (score key, ~row) of a FULL list, else 0
    float thr_f;                        // score filter: scores strictly below cannot enter the final top-k
    int n;                              // valid list entries
    int min_pos;                        // slot of the smallest entry (full list only)
    int cnt;                            // queued candidates
    uint32_t b1, b2, b3;                // three best score keys this stream has seen (b1 >= b2 >= b3)
    int hpos, hend;                     // history cursor
    uint32_t next_h;                    // next masked LOCAL row (0xFFFFFFFF when exhausted)
    int n_push, n_ins, n_steps;         // profiling counters (OOV_SCORE_DEBUG bit 2)
};

// per-CTA constants of the mask tests, all as local rows (uint32)
struct MaskRows {
    uint32_t n_rows;                    // rows >= this are TMA zero fill
    uint32_t pad_row;                   // local row of global id 0 when it is masked, else 0xFFFFFFFF
    uint32_t seg_lo, seg_hi;            // local rows kept: [seg_lo, seg_hi)
};

// smallest entry of a full list by (score key, ~row): the kept set is exactly the (score desc, row asc) top-k
__device__ __forceinline__ void list_find_min(const ScoreParams& p, UserState& u) {
    const int k = p.k;
    unsigned long long mn = ~0ull;
    int pos = 0;
#pragma unroll 4
    for (int e = 0; e < k; ++e) {
        const unsigned long long x = ((unsigned long long)u.Lh[e * SC_UG] << 32) | (unsigned long long)u.Ll[e * SC_UG];
        if (x < mn) { mn = x; pos = e; }
    }
    u.thr_hi = (uint32_t)(mn >> 32); u.thr_lo = (uint32_t)mn; u.min_pos = pos;
    u.thr_f = fmaxf(u.thr_f, float_from_order_key(u.thr_hi));     // a NaN k-th score leaves the filter unchanged
}

__device__ __forceinline__ void list_insert(const ScoreParams& p, UserState& u, uint32_t hi, uint32_t lo) {
    ++u.n_ins;
    if (hi > u.b1) { u.b3 = u.b2; u.b2 = u.b1; u.b1 = hi; }
    else if (hi > u.b2) { u.b3 = u.b2; u.b2 = hi; }
    else if (hi > u.b3) u.b3 = hi;
    if (u.n < p.k) {                                              // room left: append
        u.Lh[u.n * SC_UG] = hi; u.Ll[u.n * SC_UG] = lo;
        if (++u.n == p.k) list_find_min(p, u);
        return;
    }
    u.Lh[u.min_pos * SC_UG] = hi; u.Ll[u.min_pos * SC_UG] = lo;   // replace the smallest
    list_find_min(p, u);
}

// drop list entries whose score is strictly below the shared threshold T (they cannot be in the final top-k)
__device__ __forceinline__ void list_compact(UserState& u, uint32_t T) {
    int w = 0;
    for (int e = 0; e < u.n; ++e) {
        const uint32_t h = u.Lh[e * SC_UG], l = u.Ll[e * SC_UG];
        if (h >= T) { u.Lh[w * SC_UG] = h; u.Ll[w * SC_UG] = l; ++w; }
    }
    if (w < u.n) { u.n = w; u.thr_hi = 0u; u.thr_lo = 0u; }
}

// masks (pad / segment / history cursor) + list insertion of one queued candidate
__device__ __forceinline__ void consider(const ScoreParams& p, const MaskRows& mr, UserState& u, float s, uint32_t li) {
    if (li >= mr.n_rows) return;                                  // zero-filled rows past the end of the shard
    while (u.next_h < li) {                                       // rows arrive in increasing order
        ++u.hpos;
        u.next_h = 0xFFFFFFFFu;
        if (u.hpos < u.hend) {
            const int64_t loc = (int64_t)p.hist_cols[u.hpos] - p.item_id_offset;
            u.next_h = loc < (int64_t)mr.n_rows ? (uint32_t)loc : 0xFFFFFFFFu;      // loc >= first row of this CTA > 0
        }
    }
    if (u.next_h == li || li == mr.pad_row || li < mr.seg_lo || li >= mr.seg_hi) s = -INFINITY;
    const uint32_t hi = float_order_key(s), lo = ~li;
    if (u.n < p.k || hi > u.thr_hi || (hi == u.thr_hi && lo > u.thr_lo)) list_insert(p, u, hi, lo);
}

// all 32 lanes together: every user works through its own queue
__device__ __forceinline__ void drain(const ScoreParams& p, const MaskRows& mr, UserState& u) {
    const int n = u.cnt;
    u.n_steps += __reduce_max_sync(0xffffffffu, n);
    for (int i = 0; i < n; ++i) {
        const unsigned long long ent = u.Qe[i * SC_UG];
        consider(p, mr, u, __uint_as_float((uint32_t)(ent >> 32)), (uint32_t)ent);
    }
    u.cnt = 0;
}

// Threshold sharing, one warp for its 32 users: lane l reads the values published by streams l, l + 32, ... of one
// user ([Q][P] layout, coalesced), so lane = group; T = min over the G groups of the group's best published j-th
// score (0 while some group has published nothing).  pub_groups == 1: plain max of the published k-th bests.
__device__ __forceinline__ uint32_t shared_threshold(const ScoreParams& p, int64_t user0, int P, int lane) {
    uint32_t mine = 0u;
    const int G = p.pub_groups;
    for (int uu0 = 0; uu0 < 32; uu0 += 8) {                       // 8 users x up to 5 loads in flight per lane
        if (user0 + uu0 >= p.Q) break;                            // warp-uniform
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t user = user0 + uu0 + j;
            const uint32_t* row = p.pub + (size_t)(user < p.Q ? user : user0) * P;
            uint32_t x = 0u;
#pragma unroll
            for (int i = 0; i < SC_PUB_MAXP / 32; ++i) {
                const int s = lane + 32 * i;
                const uint32_t y = s < P ? __ldcg(row + s) : 0u;
                x = y > x ? y : x;
            }
            v[j] = x;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t T;
            if (G == 1) T = __reduce_max_sync(0xffffffffu, v[j]);
            else T = __reduce_min_sync(0xffffffffu, lane < G ? v[j] : 0xFFFFFFFFu);
            if (lane == uu0 + j) mine = T;
        }
    }
    return mine;
}

// refresh after tiles 1, 2, 3, 4, 6, 8, 12, 16, 24, ...: the useful threshold moves like 1/n
__device__ __forceinline__ bool refresh_tile(int64_t it) {
    const int64_t low = it & (it - 1);                            // `it` without its lowest set bit
    return low == 0 || ((low & (low - 1)) == 0 && (low >> 1) == (it ^ low));   // 2^a or 3 * 2^a
}

// NaN-propagating maximum of a 32-column chunk
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
    float m[11];
#pragma unroll
    for (int j = 0; j < 10; ++j)
        m[j] = max3_nan(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
    m[10] = max2_nan(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float m0 = max3_nan(m[0], m[1], m[2]), m1 = max3_nan(m[3], m[4], m[5]);
    const float m2 = max3_nan(m[6], m[7], m[8]), m3 = max2_nan(m[9], m[10]);
    return max2_nan(max3_nan(m0, m1, m2), m3);
}

// one 32-column chunk of one user's scores (registers v[0..31]); `col0` = local item row of column 0.
// Called by all 32 lanes together (warp votes inside).  Candidates are queued; the queue is drained here only when
// some lane ran out of space (then the chunk is scanned again for what is left), else at the end of the tile.
__device__ __forceinline__ void process_chunk(const ScoreParams& p, const MaskRows& mr, UserState& u, const uint32_t (&v)[32], uint32_t col0) {
    float m[11];
#pragma unroll
    for (int j = 0; j < 10; ++j)
        m[j] = max3_nan(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
    m[10] = max2_nan(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float m0 = max3_nan(m[0], m[1], m[2]), m1 = max3_nan(m[3], m[4], m[5]);
    const float m2 = max3_nan(m[6], m[7], m[8]), m3 = max2_nan(m[9], m[10]);
    const float mx = max2_nan(max3_nan(m0, m1, m2), m3);
    if (!__any_sync(0xffffffffu, !(mx < u.thr_f))) return;        // NaN compares false -> scanned
    // Common case (a lane or two with one candidate): no branches.  Upper bound of this thread's candidates = 3 x
    // groups at or above the threshold; if every lane's queue has room for that, push with predicated stores.
    int hits = 0;
#pragma unroll
    for (int g = 0; g < 11; ++g) hits += (m[g] < u.thr_f) ? 0 : 1;
    if (!__any_sync(0xffffffffu, u.cnt + 3 * hits > SC_QCAP)) {
        const float thr = u.thr_f;
        // second vote per quarter of the chunk (columns 0-8, 9-17, 18-26, 27-31): usually one quarter of one lane hits
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const float mq = qq == 0 ? m0 : (qq == 1 ? m1 : (qq == 2 ? m2 : m3));
            if (!__any_sync(0xffffffffu, !(mq < thr))) continue;
#pragma unroll
            for (int j = 9 * qq; j < (qq < 3 ? 9 * qq + 9 : 32); ++j) {
                if (!(__uint_as_float(v[j]) < thr)) {
                    u.Qe[u.cnt * SC_UG] = ((unsigned long long)v[j] << 32) | (unsigned long long)(col0 + (uint32_t)j);
                    ++u.cnt; ++u.n_push;
                }
            }
        }
        return;
    }
    // Crowded chunk (the first tiles of a stream): queue what fits, drain, scan again for the rest.
    int done = -1;                                                // last column of this chunk already queued
    while (true) {
        bool more = false;
        if (!(mx < u.thr_f)) {
#pragma unroll
            for (int g = 0; g < 11; ++g) {
                if (m[g] < u.thr_f) continue;
#pragma unroll
                for (int e3 = 0; e3 < (g < 10 ? 3 : 2); ++e3) {
                    const int idx = 3 * g + e3;
                    const uint32_t bits = v[idx];
                    if (__uint_as_float(bits) < u.thr_f || idx <= done) continue;
                    if (u.cnt < SC_QCAP) {
                        u.Qe[u.cnt * SC_UG] = ((unsigned long long)bits << 32) | (unsigned long long)(col0 + (uint32_t)idx);
                        ++u.cnt; ++u.n_push;
                        done = idx;
                    } else {
                        more = true;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, more)) break;
        drain(p, mr, u);
    }
}

// MODE 0: scoring + masks + top-k lists.  MODE 1: sampled pre-pass, per-user maximum of every visited tile.
template <int MODE>
__global__ void __launch_bounds__(SC_THREADS, 1)
tc_score_topk_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmI, const ScoreParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    unsigned char* sA = smem;                                        // SC_NUT x 16 KB
    unsigned char* sB = smem + SC_NUT * SC_A_BYTES;                  // stages x 16 KB
    uint32_t* lists_hi = reinterpret_cast<uint32_t*>(sB + p.stages * SC_B_BYTES);                     // [k][SC_UG]
    const size_t list_rows = MODE == 1 ? 0 : (size_t)p.k, queue_rows = MODE == 1 ? 0 : (size_t)SC_QCAP;   // the pre-pass keeps no lists
    uint32_t* lists_lo = lists_hi + list_rows * SC_UG;                                                 // [k][SC_UG]
    unsigned long long* queues = reinterpret_cast<unsigned long long*>(lists_lo + list_rows * SC_UG);     // [SC_QCAP][SC_UG]
    uint64_t* bars = reinterpret_cast<uint64_t*>(queues + queue_rows * SC_UG);
    uint64_t* full_bar = bars;                          // [SC_MAX_STAGES]
    uint64_t* empty_bar = bars + SC_MAX_STAGES;         // [SC_MAX_STAGES]
    uint64_t* a_full = bars + 2 * SC_MAX_STAGES;        // [1]
    uint64_t* acc_full = a_full + 1;                    // [SC_NUT]
    uint64_t* acc_empty = acc_full + SC_NUT;            // [SC_NUT]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SC_NUT);

    if (MODE == 0 && p.choice != nullptr && *p.choice == 1u) return;  // the column-split main pass was chosen (uniform exit)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.y * SC_UG;
    const int64_t q_left = p.Q - q0;
    const int n_ut = q_left >= SC_UG ? SC_NUT : (int)((q_left + SC_BM - 1) / SC_BM);   // valid 128-user tiles
    // contiguous range of visit indices of this CTA; visit i is item tile tile_begin + i * stride
    const int64_t stride = MODE == 1 ? p.tile_stride : 1;
    const int64_t t0 = p.n_visit * blockIdx.x / gridDim.x;
    const int64_t t1 = p.n_visit * (blockIdx.x + 1) / gridDim.x;

    if (warp == SC_W_TMA && lane == 0) { tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmI); }
    if (warp == SC_W_MMA && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], n_ut > 2 ? 2u : 1u); }
        mbar_init(a_full, 1);
        for (int a = 0; a < SC_NUT; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == SC_W_ALLOC) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The single-thread roles (TMA producer, MMA issuers) run with all 32 lanes of their warp in uniform control flow and
    // predicate only the TMA / MMA / commit instructions on one elected lane: under a divergent `if (lane == 0)` ptxas
    // cannot keep descriptors in uniform registers and wraps every UTCHMMA / UTMALDG in a vote loop (ELECT + R2UR.BROADCAST
    // + BRA.U.ANY, ~14 dependent instructions each — 125-200 cycles per MMA next to busy epilogue warps).
    if (warp == SC_W_TMA) {
        const bool leader = elect_one();
        if (leader) {
            mbar_arrive_expect_tx(a_full, (uint32_t)(n_ut * SC_A_BYTES));
            for (int ut = 0; ut < n_ut; ++ut) tma_load_2d(sA + ut * SC_A_BYTES, &tmU, a_full, 0, (int)(q0 + ut * SC_BM));
            // The ring holds only a few 16 KB tiles (the lists take the shared memory), less than HBM latency x the
            // rate the MMAs consume them: tiles are requested into L2 well ahead, so the ring loads are L2 hits.
            for (int64_t t = t0; t < t1 && t < t0 + SC_L2_AHEAD; ++t)
                tma_prefetch_l2_2d(&tmI, 0, (int)((p.tile_begin + t * stride) * SC_BN));
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int64_t t = t0; t < t1; ++t) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) {
                if (t + SC_L2_AHEAD < t1) tma_prefetch_l2_2d(&tmI, 0, (int)((p.tile_begin + (t + SC_L2_AHEAD) * stride) * SC_BN));
                mbar_arrive_expect_tx(&full_bar[stage], SC_B_BYTES);
                tma_load_2d(sB + stage * SC_B_BYTES, &tmI, &full_bar[stage], 0, (int)((p.tile_begin + t * stride) * SC_BN));
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= SC_W_MMA) {
        // Two issuer warps, each serving two user tiles in turn.  A single issuer for the four accumulators was the
        // bottleneck of the whole kernel: every hand-over costs it two barrier reads (~90-150 cycles each) on top of the
        // issue itself, four times per item tile, in series — more than the 4 x 256 cycles the MMAs take.  The waits
        // park the thread (try_wait), they do not poll.  (One issuer per user tile is no faster and its 672 threads
        // leave the epilogue 80 registers instead of 96.  Splitting the accumulators into 64-column halves to overlap
        // drain and MMA was tried and lost: N = 64 MMAs are shared-memory bound in SS mode.)
        const int iw = warp - SC_W_MMA;
        const bool leader = elect_one();
        if (2 * iw < n_ut) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(SC_BM, SC_BN);
            mbar_wait(a_full, 0);
            const int n_mine = n_ut - 2 * iw >= 2 ? 2 : 1;
            int stage = 0; uint32_t phase = 0, acc_phase = 0;
            for (int64_t t = t0; t < t1; ++t) {
                for (int j = 0; j < n_mine; ++j) {
                    const int ut = 2 * iw + j;
                    mbar_wait(&acc_empty[ut], acc_phase ^ 1);                     // epilogue has drained the accumulator
                    if (j == 0) mbar_wait(&full_bar[stage], phase);               // item tile has landed
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_desc(smem_u32(sA + ut * SC_A_BYTES));
                    const uint64_t bdesc = make_sw128_desc(smem_u32(sB + stage * SC_B_BYTES));
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ut * SC_BN);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
                        tc_commit(&acc_full[ut]);
                        if (j == n_mine - 1) tc_commit(&empty_bar[stage]);        // one arrival per issuer for this stage
                    }
                    __syncwarp();
                }
                acc_phase ^= 1;
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < SC_EPI_WARPS && MODE == 1) {
        // pre-pass epilogue: the user's maximum score of every visited tile, no masks, no divergence
        const int ut = (warp - SC_EPI_WARP0) >> 2;
        const int q = warp & 3;
        const int64_t user = q0 + ut * SC_BM + q * 32 + lane;
        if (ut < n_ut) {
            uint32_t acc_phase = 0;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ut * SC_BN);
            uint32_t* dst = p.tile_max + (size_t)(user < p.Q ? user : 0) * p.n_visit;
            for (int64_t t = t0; t < t1; ++t) {
                mbar_wait(&acc_full[ut], acc_phase);
                tc_fence_after();
                uint32_t v[32];
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < SC_BN / 32; ++c) {
                    tc_ld_32x32(t_lane + (uint32_t)(c * 32), v);
                    tc_wait_ld();
                    if (c == SC_BN / 32 - 1) {                         // the accumulator is in registers: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[ut]);
                    }
                    float m[11];
#pragma unroll
                    for (int j = 0; j < 10; ++j)
                        m[j] = max3_nan(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
                    m[10] = max2_nan(__uint_as_float(v[30]), __uint_as_float(v[31]));
                    const float m0 = max3_nan(m[0], m[1], m[2]), m1 = max3_nan(m[3], m[4], m[5]);
                    const float m2 = max3_nan(m[6], m[7], m[8]), m3 = max3_nan(m[9], m[10], mx);
                    mx = max2_nan(max3_nan(m0, m1, m2), m3);
                }
                if (user < p.Q) dst[t] = float_order_key(mx);
                acc_phase ^= 1;
            }
        }
    } else if (warp < SC_EPI_WARPS) {
        const int ut = (warp - SC_EPI_WARP0) >> 2;                    // which user tile / accumulator
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        const int u_local = ut * SC_BM + q * 32 + lane;               // column of `lists`
        const int64_t user = q0 + u_local;
        const bool user_ok = user < p.Q && !(p.debug & 1);
        const int k = p.k;
        UserState u;
        u.Lh = lists_hi + u_local;
        u.Ll = lists_lo + u_local;
        u.Qe = queues + u_local;
        u.thr_hi = 0u; u.thr_lo = 0u; u.n = 0; u.min_pos = 0; u.cnt = 0;
        u.b1 = u.b2 = u.b3 = 0u;
        u.n_push = u.n_ins = u.n_steps = 0;
        u.thr_f = user_ok ? -INFINITY : INFINITY;                     // padding lanes never take a candidate
        if (user_ok && p.thr_init != nullptr) {
            const uint32_t T = p.thr_init[user];
            if (T != 0u) u.thr_f = float_from_order_key(T);           // NaN (k NaN tiles): nothing is filtered
        }
        const bool share = p.share && gridDim.x > 1 && !(p.debug & 3);
        u.hpos = 0; u.hend = 0; u.next_h = 0xFFFFFFFFu;
        // mask tests on local rows (uint32): [0, N) holds global ids item_id_offset ..
        MaskRows mr;
        mr.n_rows = (uint32_t)p.N;
        mr.pad_row = (p.mask_pad && p.item_id_offset <= 0 && -p.item_id_offset < p.N) ? (uint32_t)(-p.item_id_offset) : 0xFFFFFFFFu;
        {
            const int64_t lo = p.seg_lo - p.item_id_offset, hi = p.seg_hi - p.item_id_offset;
            mr.seg_lo = lo <= 0 ? 0u : (lo >= p.N ? (uint32_t)p.N : (uint32_t)lo);
            mr.seg_hi = hi <= 0 ? 0u : (hi >= p.N ? (uint32_t)p.N : (uint32_t)hi);
        }
        uint32_t* pub_user = share ? p.pub + (size_t)(user_ok ? user : 0) * gridDim.x + blockIdx.x : nullptr;
        uint32_t last_pub = 0u;

        if (ut < n_ut) {
            if (p.hist_rowptr != nullptr && user_ok) {
                int lo = p.hist_rowptr[user];
                u.hend = p.hist_rowptr[user + 1];
                const int64_t first_gid = (p.tile_begin + t0) * SC_BN + p.item_id_offset;
                int hi = u.hend;
                while (lo < hi) {                                     // lower_bound(first id of this CTA's range)
                    const int mid = (lo + hi) >> 1;
                    if ((int64_t)p.hist_cols[mid] < first_gid) lo = mid + 1; else hi = mid;
                }
                u.hpos = lo;
                if (lo < u.hend) {
                    const int64_t loc = (int64_t)p.hist_cols[lo] - p.item_id_offset;
                    u.next_h = loc < p.N ? (uint32_t)loc : 0xFFFFFFFFu;
                }
            }
            uint32_t acc_phase = 0;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ut * SC_BN);
            for (int64_t t = t0; t < t1; ++t) {
                const int64_t it = t - t0;
                if (share && it > 0 && refresh_tile(it)) {
                    const uint32_t T = shared_threshold(p, user - lane, (int)gridDim.x, lane);
                    if (T != 0u && user_ok) {
                        u.thr_f = fmaxf(u.thr_f, float_from_order_key(T));
                        if (T != 0xFFFFFFFFu && T > u.thr_hi && !(p.debug & 8)) list_compact(u, T);      // most of a short stream's list is below T
                    }
                }
                mbar_wait(&acc_full[ut], acc_phase);
                tc_fence_after();
                const uint32_t row0 = (uint32_t)((p.tile_begin + t) * SC_BN);   // first local item row of the tile
                uint32_t v[32];
                if (p.thr_init != nullptr && !(p.debug & 64)) {
                    // With the pre-pass threshold most tiles hold nothing for these 32 users: scan the four chunks first
                    // (load, max tree, one vote each — a tcgen05.ld costs ~22 cycles) and hand the accumulator back at once
                    // when no chunk has a hit, so the next tile's MMAs start ~1000 cycles earlier; chunks with hits are
                    // loaded again and processed, the accumulator goes back after the last of them is in registers.
                    unsigned hitmask = 0u;
#pragma unroll 1
                    for (int c = 0; c < SC_BN / 32; ++c) {
                        tc_ld_32x32(t_lane + (uint32_t)(c * 32), v);
                        tc_wait_ld();
                        const float mx = chunk_max(v);
                        if (__any_sync(0xffffffffu, !(mx < u.thr_f))) hitmask |= 1u << c;
                    }
                    if (hitmask == 0u) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[ut]);
                    }
                    while (hitmask != 0u) {
                        const int c = __ffs(hitmask) - 1;
                        hitmask &= hitmask - 1u;
                        tc_ld_32x32(t_lane + (uint32_t)(c * 32), v);
                        tc_wait_ld();
                        if (hitmask == 0u) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[ut]);
                        }
                        process_chunk(p, mr, u, v, row0 + (uint32_t)(c * 32));
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < SC_BN / 32; ++c) {
                        tc_ld_32x32(t_lane + (uint32_t)(c * 32), v);
                        tc_wait_ld();
                        if (c == SC_BN / 32 - 1) {
                            // the whole accumulator has been read: hand it back to the MMA warp before the last chunk is processed
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[ut]);
                        }
                        process_chunk(p, mr, u, v, row0 + (uint32_t)(c * 32));
                    }
                }
                if (__any_sync(0xffffffffu, u.cnt > 0)) drain(p, mr, u);
                if (user_ok && share) {                               // this stream's j-th best so far
                    const uint32_t jb = p.pub_j == 1 ? u.b1 : (p.pub_j == 2 ? u.b2 : (p.pub_j == 3 ? u.b3 : u.thr_hi));
                    if (jb > last_pub) { last_pub = jb; __stcg(pub_user, jb); }
                }
                acc_phase ^= 1;
            }
        }
        if ((p.debug & 4) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0 && (warp & 3) == 0) {
            const int tp = __reduce_add_sync(0xffffffffu, u.n_push), ti = __reduce_add_sync(0xffffffffu, u.n_ins);
            const int mp = __reduce_max_sync(0xffffffffu, u.n_push);
            if (lane == 0) printf("cta %d warp %d: tiles %lld pushes %d (max lane %d) inserts %d drain-steps %d\n", (int)blockIdx.x, warp, (long long)(t1 - t0), tp, mp, ti, u.n_steps);
        }
        if (user_ok) {
            unsigned long long* dst = p.partial + ((size_t)blockIdx.x * p.Q + user) * k;
            for (int e = 0; e < u.n; ++e)
                dst[e] = ((unsigned long long)u.Lh[e * SC_UG] << 32) | (unsigned long long)u.Ll[e * SC_UG];
            p.partial_n[(size_t)blockIdx.x * p.Q + user] = (uint8_t)u.n;
        } else if (user < p.Q) {
            p.partial_n[(size_t)blockIdx.x * p.Q + user] = 0;                     // OOV_SCORE_DEBUG bit 0
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SC_W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Main pass behind a pre-pass threshold ("column-split" epilogue).  With thr_init from the sampled pre-pass a user sees
// a handful of candidates per CTA (pass rate ~1e-5), so the kernel is the speed of its scan: load, max tree, one vote.
// In tc_score_topk_kernel<0> a thread owns one USER: the four warps of a user tile read the tile's four 32-column
// chunks one after the other and the accumulator goes back to the MMA warp only then — drain (4 x ~400 cycles), hand-
// over, four MMAs and hand-over are one serial chain per accumulator (~2200 cycles per item tile against 1024 cycles of
// MMA time), and a thread that finds a candidate (masks, list insertion) is late for everything that follows.
// Here a thread owns one (TMEM lane, 32-column chunk) pair of ALL four accumulators: scan warp w reads lane quarter
// w & 3, chunk w >> 2, so the sixteen warps empty an accumulator with ONE tcgen05.ld each and hand it back at once.
// The per-user state moves out of the registers: thresholds, lists and history ranges live in shared memory.  A lane
// whose chunk holds a value at or above its user's threshold dumps the 32 scores as they are into its warp's own ring
// (single producer, sequence-stamped slots: eight 16-byte stores and one release store) and goes on; FOUR collector
// warps — one per TMEM lane quarter, so every user has exactly one writer — pick the chunks up, find the candidates with
// one vote, apply the masks (pad / segment / history) and maintain the lists (lane e holds list entry e: minimum search
// by warp reduction), raising the user's threshold when a list is full.  The history entries that fall inside the CTA's
// item range are located once per user at kernel start (two binary searches), so the collector's history test is an
// almost always empty scan.  One MMA issuer warp per accumulator (wait + descriptors + 4 MMAs + 2 commits are ~45
// dependent instructions: with two accumulators per issuer the issuers' instruction stream set the period).
// Same candidate set, same (score desc, row asc) order and the same per-CTA output lists as MODE 0: merge_keys_kernel is
// unchanged.  Measured on B200, Q = 1024, N = 10 M, D = 64, k = 20: main pass 1.39 -> 1.09 ms (all four launches 1.48 ->
// 1.22 ms).  What bounds it now is a scan warp's own chain: four steps per item tile of tcgen05.ld latency (~200 cycles
// with sixteen readers) + ~45 instructions, ~500 cycles each; two register buffers per warp would hide the load but
// need ~96 registers x 25 warps (tried with 16-column halves instead: slower, twice the votes and waits per chunk).
#ifndef OOV_SC2_NCOL
#define OOV_SC2_NCOL 4
#endif
constexpr int SC2_NCOL = OOV_SC2_NCOL;                      // collector warps: lane quarters congruent modulo SC2_NCOL (disjoint users)
constexpr int SC2_RPC = SC_EPI_WARPS / SC2_NCOL;            // rings per collector
constexpr int SC2_THREADS = (SC_EPI_WARPS + 1 + SC_NUT + SC2_NCOL) * 32;   // 800: 16 scan warps, TMA, 4 MMA issuers, 4 collectors
constexpr int SC2_W_COLLECT = SC_EPI_WARPS + 1 + SC_NUT;    // 21 .. 24
constexpr int SC2_RS = 4;                                   // hit-chunk slots per scan warp (single producer ring)
constexpr int SC2_HCAP = 4;                                 // masked local rows per user that lie in the CTA's tiles (more: CSR search)
constexpr int SC2_POOL = SC_UG * SC2_HCAP;
constexpr int SC2_SLOT_U4 = 9;                              // a slot: 32 scores (128 B) + {user column, first local row, sequence, -}
constexpr int SC2_RING_BYTES = SC_EPI_WARPS * SC2_RS * SC2_SLOT_U4 * 16;     // 9 KB

// -DOOV_SCORE_TRACE: clock-stamped events of CTA (0, 0) into p.pub (unused by this kernel): per warp 2048 64-bit slots,
// slot = clock << 24 | event << 20 | ut << 16 | tile.  scripts/trace_score.py prints the timeline.
#ifdef OOV_SCORE_TRACE
#define STRACE(ev, ut_, t_)                                                                                       \
    do {                                                                                                          \
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (t_) - t0 < 96 && tr_n < 2048) {                   \
            reinterpret_cast<unsigned long long*>(p.pub)[(size_t)warp * 2048 + tr_n++] =                          \
                ((unsigned long long)clock64() << 24) | ((unsigned long long)(ev) << 20) | ((unsigned long long)(ut_) << 16) | \
                (unsigned long long)(((t_) - t0) & 0xffff);                                                         \
        }                                                                                                         \
    } while (0)
#else
#define STRACE(ev, ut_, t_) do {} while (0)
#endif

__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");     // (measured: the fence costs nothing here)
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
        x = y < x ? y : x;
    }
    return x;
}

__global__ void __launch_bounds__(SC2_THREADS, 1)
tc_score_main2_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmI, const ScoreParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                                        // SC_NUT x 16 KB
    unsigned char* sB = smem + SC_NUT * SC_A_BYTES;                  // stages x 16 KB
    uint32_t* lists_hi = reinterpret_cast<uint32_t*>(sB + p.stages * SC_B_BYTES);      // [k][SC_UG]
    uint32_t* lists_lo = lists_hi + (size_t)p.k * SC_UG;                                // [k][SC_UG]
    uint4* ring = reinterpret_cast<uint4*>(lists_lo + (size_t)p.k * SC_UG);             // [16 warps][SC2_RS] hit-chunk slots
    float* thr_s = reinterpret_cast<float*>(ring + SC2_RING_BYTES / 16);                // [SC_UG] score filter per user
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(thr_s + SC_UG);                       // [SC_UG] valid list entries
    int32_t* h_lo = reinterpret_cast<int32_t*>(cnt_s + SC_UG);                          // [SC_UG] the user's masked rows inside this CTA's
    int32_t* h_hi = h_lo + SC_UG;                                                        //         tiles: pool[h_lo .. h_hi); h_hi < 0: not pooled
    uint32_t* pool = reinterpret_cast<uint32_t*>(h_hi + SC_UG);                          // [SC2_POOL] local rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(pool + SC2_POOL);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + SC_MAX_STAGES;
    uint64_t* a_full = bars + 2 * SC_MAX_STAGES;
    uint64_t* acc_full = a_full + 1;
    uint64_t* acc_empty = acc_full + SC_NUT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SC_NUT);
    uint32_t* done_cnt = tmem_slot + 1;
    uint32_t* head_s = tmem_slot + 4;                                                    // [16] slots the collector has consumed, per scan warp

    if (p.choice != nullptr && *p.choice != 1u) return;               // the thread-per-user main pass was chosen (uniform exit)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.y * SC_UG;
    const int64_t q_left = p.Q - q0;
    const int n_ut = q_left >= SC_UG ? SC_NUT : (int)((q_left + SC_BM - 1) / SC_BM);
    // Item tiles are dealt to the CTAs of a user group round-robin (visit i of this CTA is tile tile_begin + blockIdx.x +
    // i * gridDim.x): candidates cluster where the scores are large (e.g. the OOV half of the table), contiguous ranges
    // left half of the CTAs with 40x the hits of the other half and the kernel as slow as its busiest CTA.
    const int64_t t0 = 0;
    const int64_t t1 = (int64_t)blockIdx.x < p.n_visit ? (p.n_visit - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
#define SC2_TILE(t_) (p.tile_begin + (int64_t)blockIdx.x + (t_) * (int64_t)gridDim.x)
    const int k = p.k;
    [[maybe_unused]] int tr_n = 0;

    if (warp == SC_W_TMA && lane == 0) { tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmI); }
    if (warp == SC_W_MMA && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)n_ut); }
        mbar_init(a_full, 1);
        for (int a = 0; a < SC_NUT; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], SC_EPI_WARPS); }
        fence_barrier_init();
        *done_cnt = 0u;
        for (int w = 0; w < SC_EPI_WARPS; ++w) head_s[w] = 0u;
    }
    if (warp == SC_W_ALLOC) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < SC2_RING_BYTES / 16; i += SC2_THREADS) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < SC_UG) {
        // per-user set-up by thread = user column: threshold, empty list, masked rows inside this CTA's tiles
        const int ul = threadIdx.x;
        const int64_t user = q0 + ul;
        const bool ok = user < p.Q;
        float thr = ok ? -INFINITY : INFINITY;
        if (ok && p.thr_init != nullptr) {
            const uint32_t T = p.thr_init[user];
            if (T != 0u) thr = float_from_order_key(T);               // NaN (k NaN tiles): nothing is filtered
        }
        thr_s[ul] = thr;
        cnt_s[ul] = 0u;
        int lo = 0, hi = 0;
        if (ok && p.hist_rowptr != nullptr) {
            // One walk over the user's history, eight independent loads in flight: the masked rows that lie in this CTA's
            // tiles go to the user's SC2_HCAP pool slots, so the collector tests a candidate against a few rows in shared
            // memory; a user with more of them (very long histories) is searched in the CSR row instead.
            const int b = p.hist_rowptr[user], e = p.hist_rowptr[user + 1];
            const uint32_t gx = gridDim.x, x = blockIdx.x;
            int cnt = 0;
            for (int j0 = b; j0 < e; j0 += 8) {
                int32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = j0 + i < e ? __ldg(p.hist_cols + j0 + i) : 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (j0 + i >= e) continue;
                    const int64_t loc = (int64_t)v[i] - p.item_id_offset;
                    if (loc < 0 || loc >= p.N) continue;
                    const int64_t vt = (loc >> 7) - p.tile_begin;     // SC_BN = 128 rows per tile
                    if (vt < 0 || vt >= p.n_visit || (uint32_t)vt % gx != x) continue;
                    if (cnt < SC2_HCAP) pool[ul * SC2_HCAP + cnt] = (uint32_t)loc;
                    ++cnt;
                }
            }
            if (cnt > SC2_HCAP) hi = -1;
            else { lo = ul * SC2_HCAP; hi = lo + cnt; }
        }
        h_lo[ul] = lo; h_hi[ul] = hi;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == SC_W_TMA) {
        const bool leader = elect_one();
        if (leader) {
            mbar_arrive_expect_tx(a_full, (uint32_t)(n_ut * SC_A_BYTES));
            for (int ut = 0; ut < n_ut; ++ut) tma_load_2d(sA + ut * SC_A_BYTES, &tmU, a_full, 0, (int)(q0 + ut * SC_BM));
            for (int64_t t = t0; t < t1 && t < t0 + SC_L2_AHEAD; ++t)
                tma_prefetch_l2_2d(&tmI, 0, (int)(SC2_TILE(t) * SC_BN));
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int64_t t = t0; t < t1; ++t) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            STRACE(5, 0, t);
            if (leader) {
                if (t + SC_L2_AHEAD < t1) tma_prefetch_l2_2d(&tmI, 0, (int)(SC2_TILE(t + SC_L2_AHEAD) * SC_BN));
                mbar_arrive_expect_tx(&full_bar[stage], SC_B_BYTES);
                tma_load_2d(sB + stage * SC_B_BYTES, &tmI, &full_bar[stage], 0, (int)(SC2_TILE(t) * SC_BN));
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= SC_W_MMA && warp < SC_W_MMA + SC_NUT) {
        // One issuer warp per accumulator: wait, descriptors, four MMAs and two commits are ~45 dependent instructions
        // (~450 cycles next to the scan warps of the same scheduler) — with two accumulators per issuer the issuers'
        // instruction stream, not the tensor pipe (4 x 4 x 64 cycles per item tile), set the period of the kernel.
        const int ut = warp - SC_W_MMA;
        const bool leader = elect_one();
        if (ut < n_ut) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(SC_BM, SC_BN);
            mbar_wait(a_full, 0);
            const uint64_t adesc = make_sw128_desc(smem_u32(sA + ut * SC_A_BYTES));
            const uint32_t d_tmem = tmem_base + (uint32_t)(ut * SC_BN);
            int stage = 0; uint32_t phase = 0, acc_phase = 0;
            for (int64_t t = t0; t < t1; ++t) {
                mbar_wait(&acc_empty[ut], acc_phase ^ 1);
                STRACE(0, ut, t);
                mbar_wait(&full_bar[stage], phase);
                STRACE(1, ut, t);
                tc_fence_after();
                const uint64_t bdesc = make_sw128_desc(smem_u32(sB + stage * SC_B_BYTES));
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
                    tc_commit(&acc_full[ut]);
                    tc_commit(&empty_bar[stage]);                     // one arrival per issuer for this stage
                }
                __syncwarp();
                STRACE(2, ut, t);
                acc_phase ^= 1;
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < SC_EPI_WARPS) {
        // ===================== scan warps: lane quarter q, column chunk c of every accumulator =====================
        const int q = warp & 3, c = warp >> 2;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
        const int ul0 = q * 32 + lane;                                // user column of accumulator 0; + SC_BM per accumulator
        uint4* my_ring = ring + warp * (SC2_RS * SC2_SLOT_U4);
        uint32_t my_tail = 0;                                         // warp-uniform
        uint32_t acc_phase = 0;
        for (int64_t t = t0; t < t1; ++t) {
            const uint32_t col0 = (uint32_t)(SC2_TILE(t) * SC_BN) + (uint32_t)(c * 32);    // local item row of v[0]
#pragma unroll 1
            for (int ut = 0; ut < n_ut; ++ut) {
                const float thr = thr_s[ut * SC_BM + ul0];
                mbar_wait(&acc_full[ut], acc_phase);
                STRACE(3, ut, t);
                tc_fence_after();
                uint32_t v[32];
                tc_ld_32x32(t_lane + (uint32_t)(ut * SC_BN), v);
                tc_wait_ld();
                STRACE(4, ut, t);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ut]);           // this warp's share of the accumulator is in registers
                float m[11];
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    m[j] = max3_nan(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
                m[10] = max2_nan(__uint_as_float(v[30]), __uint_as_float(v[31]));
                const float m0 = max3_nan(m[0], m[1], m[2]), m1 = max3_nan(m[3], m[4], m[5]);
                const float m2 = max3_nan(m[6], m[7], m[8]), m3 = max2_nan(m[9], m[10]);
                const float mx = max2_nan(max3_nan(m0, m1, m2), m3);
                uint32_t hm = __ballot_sync(0xffffffffu, !(mx < thr));                 // NaN compares false -> a hit
                // Hit (a few per item tile and CTA): the lane hands its 32 scores to the collector as they are — eight
                // 16-byte stores into this warp's own ring and one release store; nothing is scanned or masked here, a
                // late scan warp is late for every accumulator that follows.
                while (hm != 0u) {
                    uint32_t take = hm;
                    int nh = __popc(take);
                    while (nh > SC2_RS) { take &= ~(0x80000000u >> __clz(take)); --nh; }       // start-up burst: SC2_RS lanes at a time
                    uint32_t polls = 0;
                    while (my_tail + (uint32_t)nh - lds_volatile(head_s + warp) > (uint32_t)SC2_RS) {   // warp-uniform
                        __nanosleep(64);
                        if (++polls > 20000000u) __trap();
                    }
                    if ((take >> lane) & 1u) {
                        const uint32_t seq = my_tail + (uint32_t)__popc(take & ((1u << lane) - 1u));
                        uint4* d = my_ring + (seq & (SC2_RS - 1)) * SC2_SLOT_U4;
#pragma unroll
                        for (int j = 0; j < 8; ++j) d[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        uint32_t* meta = reinterpret_cast<uint32_t*>(d + 8);
                        meta[0] = (uint32_t)(ut * SC_BM + ul0);
                        meta[1] = col0;
                        sts_release(meta + 2, seq / SC2_RS + 1u);
                    }
                    my_tail += (uint32_t)nh;
                    hm &= ~take;
                }
            }
            acc_phase ^= 1;
        }
        __syncwarp();
        if (lane == 0) {
            if (p.debug & 32) atomicAdd(p.pub + blockIdx.y * gridDim.x + blockIdx.x, my_tail);      // OOV_SCORE_DEBUG=32: hit chunks per CTA
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            atomicAdd(done_cnt, 1u);
        }
    } else if (warp >= SC2_W_COLLECT) {
        // ===================== collectors: masks + list maintenance =====================
        // Collector cq serves the rings of the scan warps w with w % SC2_NCOL == cq, i.e. of the lane quarters congruent to
        // cq: the users of different collectors are disjoint, so every list has one writer.  One hit chunk at a time, warp-wide:
        // lane j holds score j of the chunk (one vote finds the candidates), lane e holds list entry e (minimum search by
        // warp reduction).
        const int cq = warp - SC2_W_COLLECT;
        const uint32_t n_rows = (uint32_t)p.N;
        const uint32_t pad_row = (p.mask_pad && p.item_id_offset <= 0 && -p.item_id_offset < p.N) ? (uint32_t)(-p.item_id_offset) : 0xFFFFFFFFu;
        uint32_t seg_lo, seg_hi;
        {
            const int64_t lo = p.seg_lo - p.item_id_offset, hi = p.seg_hi - p.item_id_offset;
            seg_lo = lo <= 0 ? 0u : (lo >= p.N ? (uint32_t)p.N : (uint32_t)lo);
            seg_hi = hi <= 0 ? 0u : (hi >= p.N ? (uint32_t)p.N : (uint32_t)hi);
        }
        const int src_warp = cq + SC2_NCOL * (lane & (SC2_RPC - 1));  // lanes 0 .. SC2_RPC - 1 poll one ring each
        const uint4* src_ring = ring + src_warp * (SC2_RS * SC2_SLOT_U4);
        uint32_t my_head = 0;                                         // of ring `src_warp` (lanes < SC2_RPC)
        uint32_t idle = 0;
        while (true) {
            const uint32_t* my_meta = reinterpret_cast<const uint32_t*>(src_ring + (my_head & (SC2_RS - 1)) * SC2_SLOT_U4 + 8);
            const bool ready = lane < SC2_RPC && lds_acquire(my_meta + 2) == my_head / SC2_RS + 1u;
            const uint32_t rm = __ballot_sync(0xffffffffu, ready);
            if (rm == 0u) {
                if (lds_acquire(done_cnt) == (uint32_t)SC_EPI_WARPS) {
                    // every producer has finished and its stores are visible: one more look at the rings before leaving
                    const bool late = lane < SC2_RPC && lds_acquire(my_meta + 2) == my_head / SC2_RS + 1u;
                    if (!__any_sync(0xffffffffu, late)) break;
                    continue;
                }
                __nanosleep(idle < 8u ? 32 : 128);
                ++idle;
                continue;
            }
            idle = 0;
            const int r = __ffs(rm) - 1;                              // ring (lane) served now
            const uint32_t head_r = __shfl_sync(0xffffffffu, my_head, r);
            const int w_r = cq + SC2_NCOL * r;
            const uint4* slot = ring + w_r * (SC2_RS * SC2_SLOT_U4) + (head_r & (SC2_RS - 1)) * SC2_SLOT_U4;
            const uint32_t bits_l = reinterpret_cast<const uint32_t*>(slot)[lane];       // score `lane` of the chunk
            const uint32_t ul = reinterpret_cast<const uint32_t*>(slot + 8)[0];
            const uint32_t col0 = reinterpret_cast<const uint32_t*>(slot + 8)[1];
            if (lane == r) ++my_head;
            asm volatile("fence.acq_rel.cta;" ::: "memory");          // the reads above are done before the slot is handed back
            if (lane == r) sts_volatile(head_s + w_r, my_head);
            float thr = thr_s[ul];
            const int h0 = h_lo[ul], h1 = h_hi[ul];
            uint32_t n = cnt_s[ul];
            unsigned long long x = lane < (int)n ? (((unsigned long long)lists_hi[lane * SC_UG + ul] << 32) | (unsigned long long)lists_lo[lane * SC_UG + ul])
                                                 : ~0ull;             // list entry `lane`
            uint32_t cm = __ballot_sync(0xffffffffu, !(__uint_as_float(bits_l) < thr));
            while (cm != 0u) {
                const int j = __ffs(cm) - 1;
                cm &= cm - 1u;
                const uint32_t bits = __shfl_sync(0xffffffffu, bits_l, j);
                if (__uint_as_float(bits) < thr) continue;            // the filter rose since the vote
                const uint32_t row = col0 + (uint32_t)j;
                if (row >= n_rows) continue;                          // zero-filled rows past the end of the shard
                float sc = __uint_as_float(bits);
                bool masked = row == pad_row || row < seg_lo || row >= seg_hi;
                if (!masked && h0 < h1) {                             // rare: this user has masked rows inside the CTA's tiles
                    bool found = false;
                    for (int jj = h0 + lane; jj < h1; jj += 32) found |= pool[jj] == row;
                    masked = __any_sync(0xffffffffu, found);
                } else if (!masked && h1 < 0) {                       // not pooled: the user's CSR row in global memory
                    const int64_t user = q0 + ul;
                    const int64_t gid = (int64_t)row + p.item_id_offset;
                    bool found = false;
                    for (int jj = p.hist_rowptr[user] + lane; jj < p.hist_rowptr[user + 1]; jj += 32) found |= (int64_t)p.hist_cols[jj] == gid;
                    masked = __any_sync(0xffffffffu, found);
                }
                if (masked) sc = -INFINITY;
                const unsigned long long cand = ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)(~row);
                unsigned long long kth = 0ull;
                bool moved = false;
                if ((int)n < k) {
                    if (lane == (int)n) x = cand;
                    ++n;
                    if ((int)n == k) { kth = warp_min_u64(x); moved = true; }
                } else {
                    const unsigned long long mn = warp_min_u64(x);
                    if (cand > mn) {
                        if (x == mn) x = cand;                        // entries are distinct (the row is part of the key)
                        kth = warp_min_u64(x);
                        moved = true;
                    }
                }
                if (moved) thr = fmaxf(thr, float_from_order_key((uint32_t)(kth >> 32)));     // a NaN k-th score leaves the filter unchanged
            }
            // write the list back (only the entries that exist), the count and the filter
            if (lane < (int)n) { lists_hi[lane * SC_UG + ul] = (uint32_t)(x >> 32); lists_lo[lane * SC_UG + ul] = (uint32_t)x; }
            if (lane == 0) { cnt_s[ul] = n; thr_s[ul] = thr; }
            __syncwarp();
        }
    }

    // scan warps + collector: the lists are final
    if (warp < SC_EPI_WARPS || warp >= SC2_W_COLLECT) {
        asm volatile("bar.sync 1, %0;" ::"n"((SC_EPI_WARPS + SC2_NCOL) * 32) : "memory");
        if (warp < SC_EPI_WARPS) {
            const int ul = threadIdx.x;
            const int64_t user = q0 + ul;
            if (user < p.Q) {
                const int n = (int)cnt_s[ul];
                unsigned long long* dst = p.partial + ((size_t)blockIdx.x * p.Q + user) * k;
                for (int e = 0; e < n; ++e)
                    dst[e] = ((unsigned long long)lists_hi[e * SC_UG + ul] << 32) | (unsigned long long)lists_lo[e * SC_UG + ul];
                p.partial_n[(size_t)blockIdx.x * p.Q + user] = (uint8_t)n;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SC_W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Pre-pass threshold: one warp per user.  R = k + (masked items of the user that lie in sampled tiles); the R-th
// largest of the user's n_s tile maxima (ordered keys, radix select 4 x 8 bits over a warp-private histogram) is a
// bound no top-k score is below.  0 = no bound (fewer than R sampled tiles).
constexpr int THR_WARPS = 8;
constexpr int THR_UNR = 16;
__global__ void __launch_bounds__(THR_WARPS * 32)
score_threshold_kernel(const uint32_t* __restrict__ tile_max, int64_t n_s, int64_t Q, int k, int64_t tile_first,
                       int64_t tile_stride, int64_t item_id_offset, int64_t N, int mask_pad,
                       const int32_t* __restrict__ hist_rowptr, const int32_t* __restrict__ hist_cols,
                       uint32_t* __restrict__ thr_out, uint32_t* __restrict__ tile_hits) {
    __shared__ uint32_t hist_s[THR_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t user = (int64_t)blockIdx.x * THR_WARPS + warp;
    if (user >= Q) return;
    uint32_t* hist = hist_s[warp];
    auto sampled = [&](int64_t loc) -> bool {                         // local row inside a sampled tile?
        if (loc < 0 || loc >= N) return false;
        const int64_t d = loc / SC_BN - tile_first;
        return d >= 0 && d % tile_stride == 0 && d / tile_stride < n_s;
    };
    int h = 0;
    if (hist_rowptr != nullptr) {
        const int e = hist_rowptr[user + 1];
        for (int j = hist_rowptr[user] + lane; j < e; j += 32) h += sampled((int64_t)hist_cols[j] - item_id_offset) ? 1 : 0;
    }
    h = __reduce_add_sync(0xffffffffu, h);
    if (mask_pad && sampled(-item_id_offset)) ++h;
    int64_t remaining = (int64_t)k + h;                               // rank (1 = largest) still to find
    if (remaining > n_s) { if (lane == 0) thr_out[user] = 0u; return; }
    const uint32_t* row = tile_max + (size_t)user * n_s;
    uint32_t prefix = 0u, mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0u;
        __syncwarp();
        // THR_UNR loads in flight per lane: the pass is latency-bound (one warp walks its user's row: ~7 warps per SM)
        for (int64_t i0 = 0; i0 < n_s; i0 += 32 * THR_UNR) {
            uint32_t v[THR_UNR];
#pragma unroll
            for (int j = 0; j < THR_UNR; ++j) {
                const int64_t i = i0 + j * 32 + lane;
                v[j] = i < n_s ? __ldg(row + i) : 0u;
            }
#pragma unroll
            for (int j = 0; j < THR_UNR; ++j)
                if (i0 + j * 32 + lane < n_s && (v[j] & mask) == prefix) atomicAdd(&hist[(v[j] >> shift) & 255u], 1u);
        }
        __syncwarp();
        uint32_t c[8], lane_sum = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; lane_sum += c[j]; }
        // values in bins above this lane's eight bins
        uint32_t incl = lane_sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_down_sync(0xffffffffu, incl, d);
            if (lane + d < 32) incl += y;
        }
        uint32_t acc = incl - lane_sum;
        int found = -1;
        uint32_t rem_new = 0u;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            if (found < 0 && (int64_t)acc < remaining && (int64_t)(acc + c[j]) >= remaining) { found = lane * 8 + j; rem_new = (uint32_t)(remaining - acc); }
            acc += c[j];
        }
        const unsigned who = __ballot_sync(0xffffffffu, found >= 0);  // exactly one lane
        const int src = __ffs(who) - 1;
        found = __shfl_sync(0xffffffffu, found, src);
        rem_new = __shfl_sync(0xffffffffu, rem_new, src);
        prefix |= (uint32_t)found << shift;
        mask |= 255u << shift;
        remaining = rem_new;
        __syncwarp();
    }
    if (lane == 0) thr_out[user] = prefix;
    if (tile_hits != nullptr)                                         // how many users will see a candidate in sampled tile i
        for (int64_t i0 = 0; i0 < n_s; i0 += 32 * THR_UNR) {
            uint32_t v[THR_UNR];
#pragma unroll
            for (int j = 0; j < THR_UNR; ++j) {
                const int64_t i = i0 + j * 32 + lane;
                v[j] = i < n_s ? __ldg(row + i) : 0u;                 // 0 < any threshold key that filters (prefix == 0: no bound, every tile counts)
            }
#pragma unroll
            for (int j = 0; j < THR_UNR; ++j)
                if (i0 + j * 32 + lane < n_s && v[j] >= prefix) atomicAdd(tile_hits + i0 + j * 32 + lane, 1u);
        }
}

// Which main pass runs (device-side, so the choice can sit inside a captured graph).  Every user has exactly R sampled
// tiles at or above its threshold; what differs between tables is how those tiles COINCIDE across users.  With
// embeddings whose scores are dominated by the item (all-positive DHE rows: an item that scores high does so for every
// such user) a few dozen tiles are hit by half of all users at once, and a scan warp of the column-split pass then has to
// hand off up to 32 hit chunks in one step through a 4-slot ring — the thread-per-user pass takes such bursts in
// parallel.  Statistic: h_i = users with a candidate in sampled tile i; co-hit = sum h_i^2 / sum h_i (how many users share
// the tile of a random hit) against the mean.  Measured on the bench tables (1953 sampled tiles, mean 12): LSH 27 (2.2 x
// the mean; main pass 0.52 -> 0.45 ms with the column-split kernel), DHE 296 (25 x; 0.17 -> 0.34 ms).  One tile with a
// NaN item (a candidate of all 1024 users) adds ~45.  out[0] = 1 (column-split) while co-hit <= SC2_COHIT x mean.
constexpr unsigned long long SC2_COHIT = 12;
__global__ void score_choose_kernel(const uint32_t* __restrict__ tile_hits, int64_t n_s, uint32_t* __restrict__ out, int verbose) {
    __shared__ unsigned long long s_sq[32], s_sum[32];
    unsigned long long sq = 0ull, sum = 0ull;
    for (int64_t i = threadIdx.x; i < n_s; i += blockDim.x) { const unsigned long long v = tile_hits[i]; sq += v * v; sum += v; }
    for (int o = 16; o; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
    if ((threadIdx.x & 31) == 0) { s_sq[threadIdx.x >> 5] = sq; s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { sq += s_sq[w]; sum += s_sum[w]; }
        // co-hit <= SC2_COHIT * mean  <=>  sq * n_s <= SC2_COHIT * sum^2   (sum <= Q * R ~ 1e5, sq * n_s < 2^63)
        out[0] = (sum > 0ull && sq * (unsigned long long)n_s <= SC2_COHIT * sum * sum) ? 1u : 0u;
        out[1] = (uint32_t)(sum > 0ull ? sq / sum : 0ull); out[2] = (uint32_t)(sum > 0xFFFFFFFFull ? 0xFFFFFFFFull : sum); out[3] = (uint32_t)n_s;
        if (verbose) printf("score_choose: column-split %u (co-hit %.1f, mean %.1f over %lld sampled tiles)\n", out[0],
                            sum > 0ull ? (double)sq / (double)sum : 0.0, n_s > 0 ? (double)sum / (double)n_s : 0.0, (long long)n_s);
    }
}

static int score_stages(int k) {
    const int fixed = 1024 + SC_NUT * SC_A_BYTES + (k + SC_QCAP) * SC_UG * 8 + 512;
    int s = (SC_SMEM_MAX - fixed) / SC_B_BYTES;
    return s > SC_MAX_STAGES ? SC_MAX_STAGES : s;
}
static size_t score_smem_bytes(int k, int stages) {
    return 1024 + SC_NUT * SC_A_BYTES + (size_t)stages * SC_B_BYTES + (size_t)(k + SC_QCAP) * SC_UG * 8 + 512;
}
// column-split main pass: lists, candidate ring, per-user threshold / count / history range
static size_t score2_fixed_bytes(int k) {
    return 1024 + SC_NUT * SC_A_BYTES + (size_t)k * SC_UG * 8 + (size_t)SC2_RING_BYTES + (size_t)SC_UG * 16 + (size_t)SC2_POOL * 4 + 512;
}
static int score2_stages(int k) {
    int s = (int)((SC_SMEM_MAX - score2_fixed_bytes(k)) / SC_B_BYTES);
    return s > SC_MAX_STAGES ? SC_MAX_STAGES : s;
}
// Which main pass follows the sampled pre-pass.  The column-split kernel wins while hits are rare — every hit chunk makes
// one scan warp late — and loses to MODE 0 (every thread maintains its own user's list, 512 in parallel) when they are
// not: a user sees ~(k + masked) x stride items above the sampled threshold, whatever N is, so the hit chunks per item
// tile and CTA are ~(k + 4) x stride x 512 / tiles.  Measured on B200 (Q = 1024, k = 20): 2.4 per tile (10 M items,
// stride 16) 1.60 -> 1.34 ms; 6 per tile (1 M items, stride 4, DHE table) 0.23 -> 0.36 ms.
// OOV_SCORE_MAIN2 (read on every call; tests and profiling): 0 = never, 2 = whenever the pre-pass runs, else automatic.
static bool score2_wanted(int k, int64_t stride, int64_t n_tiles) {
    const char* e = getenv("OOV_SCORE_MAIN2");
    const int mode = e ? atoi(e) : 1;
    if (mode == 0) return false;
    if (mode == 2) return true;
    // crossover measured on LSH tables: 1.25 M rows at stride 4 (5 hit chunks per tile) 0.318 -> 0.304 ms, 2.5 M rows
    // 0.52 -> 0.45 ms; below ~8 k tiles the start-up of the 800-thread kernel and the hit rate eat the gain
    return 2 * (int64_t)(k + 4) * stride * SC_UG <= 11 * n_tiles;
}

bool score_tc_supported(int dtype, int D, int k) {
    return dtype == OOV_BF16 && D >= 8 && D <= 64 && D % 8 == 0 && k >= 1 && k <= SC_KMAX && score_stages(k) >= 2;
}

static int score_grid_x(int64_t Q, int64_t n_tiles) {
    const int64_t groups = cdiv(Q, SC_UG);
    int64_t gx = num_sms() / groups;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    if (gx > SC_PUB_MAXP) gx = SC_PUB_MAXP;
    if (gx < 1) gx = 1;
    return (int)gx;
}

static size_t score_pub_bytes(int64_t Q, int gx) { return align_up((size_t)gx * Q * 4, 256); }
static size_t score_partial_bytes(int64_t Q, int gx, int k) { return align_up((size_t)gx * Q * k * 8, 256) + align_up((size_t)gx * Q, 256); }

// sampled pre-pass: every `stride`-th full tile once the shard has enough tiles for a useful threshold
constexpr int64_t SC_PRE_MIN_TILES = 256;
static int64_t score_pre_stride(int64_t n_full_tiles) {
    if (n_full_tiles < SC_PRE_MIN_TILES) return 0;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("OOV_SCORE_STRIDE"); forced = e ? atoi(e) : 0; }   // profiling only; 1 = no pre-pass
    if (forced == 1) return 0;
    if (forced > 1) return forced;
    // The pre-pass costs ~tiles / stride, the candidates it leaves cost ~stride (a user sees ~(k + masked) x stride items
    // above its threshold whatever N is): measured optima on B200 (Q = 1024, k = 20) — 1 k tiles: 2; 8 k - 20 k tiles: 4
    // (1.25 M rows 0.343 -> 0.30 ms, 2.5 M rows 0.57 -> 0.45 ms against the old "at least 1024 sampled tiles" rule);
    // 39 k tiles: 8 (0.875 -> 0.77 ms); 78 k tiles: 16 (stride 8: 1.41 against 1.35 ms).
    if (n_full_tiles < 4096) return 2;
    if (n_full_tiles < 32768) return 4;
    if (n_full_tiles < 65536) return 8;
    return 16;
}
// most tiles any kept segment of an N-row shard can sample (the stride grows with the segment)
static int64_t score_pre_max_visits(int64_t N) {
    const int64_t nf = N / SC_BN;
    if (score_pre_stride(nf) == 0) return 0;
    int64_t best = 0;
    for (int64_t m = SC_PRE_MIN_TILES; ; m *= 2) {                    // a kept segment has m <= nf full tiles: visits = ceil(m / stride(m))
        const int64_t mm = m < nf ? m : nf;
        for (int64_t d = 0; d < 2; ++d) {                             // just below every stride boundary and at the size itself
            const int64_t x = mm - d > 0 ? mm - d : mm;
            const int64_t st = score_pre_stride(x);
            if (st > 0 && cdiv(x, st) > best) best = cdiv(x, st);
        }
        if (mm == nf) break;
    }
    for (int64_t b : {int64_t(4095), int64_t(32767), int64_t(65535)})
        if (b <= nf && cdiv(b, score_pre_stride(b)) > best) best = cdiv(b, score_pre_stride(b));
    return best;
}
static size_t score_thr_bytes(int64_t Q) { return align_up((size_t)Q * 4, 256); }
static size_t score_hits_bytes(int64_t nv) { return align_up((size_t)nv * 4 + 16, 256); }     // per sampled tile + the 4-word choice record
static size_t score_pre_bytes(int64_t Q, int64_t N) {
    const int64_t nv = score_pre_max_visits(N);
    return nv == 0 ? 0 : score_thr_bytes(Q) + align_up((size_t)Q * (size_t)nv * 4, 256) + score_hits_bytes(nv);
}

size_t score_tc_workspace(int64_t Q, int64_t N, int k) {
    const int gx = score_grid_x(Q, cdiv(N > 0 ? N : 1, SC_BN));
    return score_pub_bytes(Q, gx) + score_partial_bytes(Q, gx, k) + score_pre_bytes(Q, N);
}

int score_tc_run(const void* users, const void* items, int64_t Q, int64_t N, int D, int k, int64_t item_id_offset,
                 int mask_pad, int64_t seg_lo, int64_t seg_hi, const int32_t* hist_rowptr, const int32_t* hist_cols,
                 float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t st, const KeyOut ko) {
    // only item tiles that intersect the kept segment are visited
    int64_t lo = seg_lo - item_id_offset, hi = seg_hi > item_id_offset + N ? N : seg_hi - item_id_offset;
    if (lo < 0) lo = 0;
    if (hi > N) hi = N;
    ScoreParams p{};
    p.Q = Q; p.N = N; p.k = k; p.item_id_offset = item_id_offset; p.mask_pad = mask_pad;
    p.seg_lo = seg_lo; p.seg_hi = seg_hi; p.hist_rowptr = hist_rowptr; p.hist_cols = hist_cols;
    p.tile_begin = hi > lo ? lo / SC_BN : 0;
    p.tile_end = hi > lo ? cdiv(hi, SC_BN) : 0;
    p.stages = score_stages(k);
    { const char* dbg = getenv("OOV_SCORE_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    const int64_t n_tiles = p.tile_end - p.tile_begin;
    const int gx = score_grid_x(Q, n_tiles > 0 ? n_tiles : 1);
    const size_t need = score_pub_bytes(Q, gx) + score_partial_bytes(Q, gx, k);
    OOV_REQUIRE(workspace && workspace_bytes >= need, OOV_ERR_WORKSPACE, "oov_fullsort_topk (tcgen05): workspace %zu < %zu",
                workspace_bytes, need);
    OOV_REQUIRE(aligned(workspace, 8), OOV_ERR_ALIGN, "oov_fullsort_topk (tcgen05): workspace must be 8-byte aligned");
    p.pub = reinterpret_cast<uint32_t*>(workspace);
    p.partial = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + score_pub_bytes(Q, gx));
    p.partial_n = reinterpret_cast<uint8_t*>(p.partial) + align_up((size_t)gx * Q * k * 8, 256);
    p.pub_groups = gx < 32 ? gx : 32;                                // group = streams congruent modulo 32
    p.pub_j = (k + p.pub_groups - 1) / p.pub_groups;
    if (p.pub_j > 3) { p.pub_groups = 1; p.pub_j = k; }              // few long streams: share the plain k-th best
    p.n_visit = n_tiles; p.tile_stride = 1; p.share = 1;

    // sampled pre-pass over the FULL tiles of the kept segment (no zero-filled or segment-masked rows inside)
    const int64_t full_begin = hi > lo ? cdiv(lo, SC_BN) : 0, full_end = hi > lo ? hi / SC_BN : 0;
    const int64_t pre_stride = score_pre_stride(full_end - full_begin);
    ScoreParams pa = p;
    uint32_t* tile_hits = nullptr;
    if (pre_stride > 0) {
        pa.stages = SC_MAX_STAGES;                                    // no lists in the pre-pass: the whole shared memory is TMA ring
        pa.tile_begin = full_begin;
        pa.tile_stride = pre_stride;
        pa.n_visit = cdiv(full_end - full_begin, pre_stride);
        unsigned char* w = reinterpret_cast<unsigned char*>(workspace) + score_pub_bytes(Q, gx) + score_partial_bytes(Q, gx, k);
        uint32_t* thr = reinterpret_cast<uint32_t*>(w);
        pa.tile_max = reinterpret_cast<uint32_t*>(w + score_thr_bytes(Q));
        const size_t tm_bytes = align_up((size_t)Q * (size_t)pa.n_visit * 4, 256);
        tile_hits = reinterpret_cast<uint32_t*>(w + score_thr_bytes(Q) + tm_bytes);     // [n_visit] + 4 words (choice, max, sum, n)
        const size_t need_pre = (size_t)(w - reinterpret_cast<unsigned char*>(workspace)) + score_thr_bytes(Q) + tm_bytes + score_hits_bytes(pa.n_visit);
        OOV_REQUIRE(workspace_bytes >= need_pre, OOV_ERR_WORKSPACE, "oov_fullsort_topk (tcgen05): workspace %zu < %zu (pre-pass)",
                    workspace_bytes, need_pre);
        p.thr_init = thr;
        p.share = (p.debug & 16) ? 1 : 0;                            // the sampled threshold beats the exchanged one from the first tile
    }
    if (p.share) {
        cudaError_t ce = cudaMemsetAsync(p.pub, 0, (size_t)gx * Q * 4, st);
        OOV_REQUIRE(ce == cudaSuccess, OOV_ERR_CUDA, "cudaMemsetAsync(threshold array): %s", cudaGetErrorString(ce));
    }

    CUtensorMap tmU, tmI;
    int rc = make_tmap_bf16_2d(&tmU, users, (uint64_t)D, (uint64_t)Q, (uint64_t)D * 2, SC_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmI, items, (uint64_t)D, (uint64_t)(N > 0 ? N : 1), (uint64_t)D * 2, SC_BN);
    if (rc) return rc;
    const size_t smem = score_smem_bytes(k, p.stages);
    {   // per device / context, cheap and idempotent: set on every call (a process may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(tc_score_topk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_MAX);
        OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_score_topk_kernel<0>): %s", cudaGetErrorString(e));
        e = cudaFuncSetAttribute(tc_score_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_MAX);
        OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_score_topk_kernel<1>): %s", cudaGetErrorString(e));
    }
    const dim3 grid((unsigned)gx, (unsigned)cdiv(Q, SC_UG));
    // main pass: the column-split kernel when the table is large enough for it (score2_wanted), chosen against the
    // thread-per-user kernel ON THE DEVICE from the burstiness of the sampled hits (both are launched, one returns at once)
    const char* m2 = getenv("OOV_SCORE_MAIN2");
    const int m2_mode = m2 ? atoi(m2) : 1;
    const bool try2 = pre_stride > 0 && !p.share && !(p.debug & ~32) && score2_stages(k) >= 2 && score2_wanted(k, pre_stride, n_tiles);
    const bool adaptive = try2 && m2_mode == 1;
    if (pre_stride > 0) {
        tc_score_topk_kernel<1><<<grid, SC_THREADS, score_smem_bytes(-SC_QCAP, pa.stages), st>>>(tmU, tmI, pa);
        OOV_LAUNCH_CHECK("tc_score_topk_kernel<1> (pre-pass)");
        if (adaptive) {
            cudaError_t ce = cudaMemsetAsync(tile_hits, 0, (size_t)pa.n_visit * 4, st);
            OOV_REQUIRE(ce == cudaSuccess, OOV_ERR_CUDA, "cudaMemsetAsync(tile hits): %s", cudaGetErrorString(ce));
        }
        score_threshold_kernel<<<(unsigned)cdiv(Q, THR_WARPS), THR_WARPS * 32, 0, st>>>(
            pa.tile_max, pa.n_visit, Q, k, pa.tile_begin, pa.tile_stride, item_id_offset, N, mask_pad, hist_rowptr, hist_cols,
            const_cast<uint32_t*>(p.thr_init), adaptive ? tile_hits : nullptr);
        OOV_LAUNCH_CHECK("score_threshold_kernel");
        if (adaptive) {
            score_choose_kernel<<<1, 256, 0, st>>>(tile_hits, pa.n_visit, tile_hits + pa.n_visit, (p.debug & 32) ? 1 : 0);
            OOV_LAUNCH_CHECK("score_choose_kernel");
            p.choice = tile_hits + pa.n_visit;
        }
    }
    if (try2) {
        if (p.debug & 32) {
            cudaError_t ce = cudaMemsetAsync(p.pub, 0, (size_t)gx * cdiv(Q, SC_UG) * 4, st);
            OOV_REQUIRE(ce == cudaSuccess, OOV_ERR_CUDA, "cudaMemsetAsync(hit counters): %s", cudaGetErrorString(ce));
        }
        cudaError_t e = cudaFuncSetAttribute(tc_score_main2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_MAX);
        OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_score_main2_kernel): %s", cudaGetErrorString(e));
        ScoreParams p2 = p;
        p2.stages = score2_stages(k);
        tc_score_main2_kernel<<<grid, SC2_THREADS, score2_fixed_bytes(k) + (size_t)p2.stages * SC_B_BYTES, st>>>(tmU, tmI, p2);
        OOV_LAUNCH_CHECK("tc_score_main2_kernel");
    }
    if (!try2 || adaptive) {
        tc_score_topk_kernel<0><<<grid, SC_THREADS, smem, st>>>(tmU, tmI, p);
        OOV_LAUNCH_CHECK("tc_score_topk_kernel");
    }
    return launch_merge_keys(p.partial, p.partial_n, gx, Q, k, item_id_offset, out_scores, out_idx, st, ko);
}

}  // namespace tc
}  // namespace oov
