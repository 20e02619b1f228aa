// Full-sort scoring + masks + top-k — CUDA-core (fp32 FMA) path, plus the shard merge and the
// collector hit matrix.
//
// Restates bpr.py:151-156 / directau.py:193-198 (score = user_e @ all_item_e.T),
// inductive/evaluator.py:91-94 (pad column and history -> -inf), evaluator/collector.py:153-166
// (torch.topk + hit gather) and inductive/collector_filter.py:172-175 (item-segment filter) of the
// reference — without ever materialising the [Q, N] score matrix: every warp streams item rows,
// keeps a register-resident sorted top-k list per user (lane j holds entry j), and only items
// that beat the current k-th entry pay for the mask checks and the insertion.
// The tensor-core (tcgen05) variant lives in tc_score.cu and shares the merge kernels below.
#include "common.cuh"

namespace oov {

// ---------------------------------------------------------------- warp-resident sorted list
// KR registers per lane, entry e = r * 32 + lane, sorted descending by key64 (0 = empty).
template <int KR>
struct WarpList {
    unsigned long long e[KR];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int r = 0; r < KR; ++r) e[r] = 0ull;
    }
    // k-th best so far (entry k-1), broadcast to the warp
    __device__ __forceinline__ unsigned long long kth(int k) const {
        const int kk = k - 1;
        unsigned long long v = 0;
#pragma unroll
        for (int r = 0; r < KR; ++r)
            if ((kk >> 5) == r) v = __shfl_sync(0xffffffffu, e[r], kk & 31);
        return v;
    }
    // insert a warp-uniform key x (all lanes call); entries beyond k are harmless extra slots
    __device__ __forceinline__ void insert(unsigned long long x, int lane) {
        unsigned long long carry = x;              // value entering register r at its insertion point / lane 0
        bool placed = false;
#pragma unroll
        for (int r = 0; r < KR; ++r) {
            const unsigned m = __ballot_sync(0xffffffffu, carry > e[r]);   // suffix of lanes with smaller entries
            const unsigned long long last = __shfl_sync(0xffffffffu, e[r], 31);
            if (placed) {
                // whole register shifts up by one, lane 0 takes the previous register's last entry
                const unsigned long long up = __shfl_up_sync(0xffffffffu, e[r], 1);
                e[r] = lane == 0 ? carry : up;
                carry = last;
            } else if (m) {
                const int pos = __ffs(m) - 1;
                const unsigned long long up = __shfl_up_sync(0xffffffffu, e[r], 1);
                if (lane > pos) e[r] = up;
                else if (lane == pos) e[r] = carry;
                carry = last;
                placed = true;
            }
        }
    }
};

__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ cols, int lo, int hi, int64_t gid) {
    int l = lo, h = hi;
    while (l < h) {
        const int mid = (l + h) >> 1;
        if ((int64_t)cols[mid] < gid) l = mid + 1; else h = mid;
    }
    return l < hi && (int64_t)cols[l] == gid;
}

template <typename T> __device__ __forceinline__ void load16(const T* p, float (&x)[16]);
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + j);
        x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(p) + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(h[q]);
            x[8 * j + 2 * q] = f.x; x[8 * j + 2 * q + 1] = f.y;
        }
    }
}

constexpr int FS_THREADS = 256;
constexpr int FS_WARPS = FS_THREADS / 32;

// grid = (item CTAs, user tiles).  Each warp owns 32-item batches (one item per lane), strided over
// the grid, and TQ register lists.  Requires D % 16 == 0 and 16-byte aligned rows (checked on host).
template <typename T, int TQ, int KR>
__global__ void __launch_bounds__(FS_THREADS)
fullsort_topk_simt(const T* __restrict__ users, const T* __restrict__ items, int64_t Q, int64_t N, int D, int k,
                   int64_t item_id_offset, int mask_pad, int64_t seg_lo, int64_t seg_hi,
                   const int32_t* __restrict__ hist_rowptr, const int32_t* __restrict__ hist_cols,
                   unsigned long long* __restrict__ partial /* [gridDim.x][Q][k] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* us = reinterpret_cast<float*>(smem_raw);                                   // [TQ][D]
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(us + TQ * D);   // [FS_WARPS][TQ][KR*32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * TQ;

    for (int i = tid; i < TQ * D; i += FS_THREADS) {
        const int u = i / D, d = i - u * D;
        us[i] = (q0 + u < Q) ? to_f32<T>(users[(q0 + u) * D + d]) : 0.f;
    }
    __syncthreads();

    WarpList<KR> wl[TQ];
#pragma unroll
    for (int u = 0; u < TQ; ++u) wl[u].clear();
    unsigned long long thr[TQ];
#pragma unroll
    for (int u = 0; u < TQ; ++u) thr[u] = 0ull;
    int hlo[TQ], hhi[TQ];
#pragma unroll
    for (int u = 0; u < TQ; ++u) {
        hlo[u] = hhi[u] = 0;
        if (hist_rowptr != nullptr && q0 + u < Q) { hlo[u] = hist_rowptr[q0 + u]; hhi[u] = hist_rowptr[q0 + u + 1]; }
    }

    const int64_t n_batches = (N + 31) / 32;
    for (int64_t b = (int64_t)blockIdx.x * FS_WARPS + warp; b < n_batches; b += (int64_t)gridDim.x * FS_WARPS) {
        const int64_t li = b * 32 + lane;                     // local item row of this lane
        const bool valid = li < N;
        const T* row = items + (valid ? li : 0) * D;
        float acc[TQ];
#pragma unroll
        for (int u = 0; u < TQ; ++u) acc[u] = 0.f;
        for (int d0 = 0; d0 < D; d0 += 16) {
            float x[16];
            load16<T>(row + d0, x);
#pragma unroll
            for (int u = 0; u < TQ; ++u) {
                const float4* up = reinterpret_cast<const float4*>(us + u * D + d0);
                float a = acc[u];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = up[j];
                    a = fmaf(x[4 * j], w.x, a); a = fmaf(x[4 * j + 1], w.y, a);
                    a = fmaf(x[4 * j + 2], w.z, a); a = fmaf(x[4 * j + 3], w.w, a);
                }
                acc[u] = a;
            }
        }
        const int64_t gid = li + item_id_offset;
        // per-item masks: pad item 0 (evaluator.py:92) and the item-segment filter (collector_filter.py:172-175)
        const bool item_masked = (mask_pad && gid == 0) || gid < seg_lo || gid >= seg_hi;
#pragma unroll
        for (int u = 0; u < TQ; ++u) {
            if (q0 + u >= Q) continue;                         // warp-uniform
            unsigned long long key = valid ? make_key64(item_masked ? -INFINITY : acc[u], (uint32_t)li) : 0ull;
            unsigned m = __ballot_sync(0xffffffffu, key > thr[u]);
            while (m) {                                        // rare after warm-up
                const int src = __ffs(m) - 1;
                m &= m - 1;
                unsigned long long x = __shfl_sync(0xffffffffu, key, src);
                const int64_t xg = __shfl_sync(0xffffffffu, gid, src);
                // history (evaluator.py:93-94) is only probed for candidates that beat the threshold
                if (hhi[u] > hlo[u] && sorted_contains(hist_cols, hlo[u], hhi[u], xg))
                    x = make_key64(-INFINITY, key64_idx(x));
                if (x > thr[u]) {
                    wl[u].insert(x, lane);
                    thr[u] = wl[u].kth(k);
                }
            }
        }
    }

    // CTA-level merge of the FS_WARPS lists per user, then one partial list per (CTA, user)
#pragma unroll
    for (int u = 0; u < TQ; ++u)
#pragma unroll
        for (int r = 0; r < KR; ++r) lists[((size_t)warp * TQ + u) * (KR * 32) + r * 32 + lane] = wl[u].e[r];
    __syncthreads();
    for (int u = warp; u < TQ; u += FS_WARPS) {
        if (q0 + u >= Q) continue;
        WarpList<KR> fin;
        fin.clear();
        unsigned long long t = 0ull;
        for (int w = 0; w < FS_WARPS; ++w) {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const unsigned long long key = lists[((size_t)w * TQ + u) * (KR * 32) + r * 32 + lane];
                unsigned m = __ballot_sync(0xffffffffu, key > t);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long x = __shfl_sync(0xffffffffu, key, src);
                    if (x > t) { fin.insert(x, lane); t = fin.kth(k); }
                }
            }
        }
        unsigned long long* dst = partial + ((size_t)blockIdx.x * Q + (q0 + u)) * k;
#pragma unroll
        for (int r = 0; r < KR; ++r)
            if (r * 32 + lane < k) dst[r * 32 + lane] = fin.e[r];
    }
}

// one warp per user: merge P partial lists of up to k key64 entries; decode to (score, global id).
// `partial_n` (optional): [P][Q] valid entries per list — the tensor-core kernel's lists are mostly short, and only
// the valid entries are read (and were written).
template <int KR>
__global__ void __launch_bounds__(256)
merge_keys_kernel(const unsigned long long* __restrict__ partial, const uint8_t* __restrict__ partial_n, int P, int64_t Q, int k,
                  int64_t item_id_offset, float* __restrict__ out_scores, int64_t* __restrict__ out_idx, const KeyOut ko) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    WarpList<KR> fin;
    fin.clear();
    unsigned long long t = 0ull;
    constexpr int MU = 8;                                   // loads in flight per lane: the loop is latency-bound otherwise
    if (partial_n != nullptr) {
        // lane l owns lists l, l + 32, ...: walk them entry by entry, MU lists at a time
        for (int p0 = 0; p0 < P; p0 += 32 * MU) {
            int n[MU], nmax = 0;
#pragma unroll
            for (int u = 0; u < MU; ++u) {
                const int p = p0 + u * 32 + lane;
                n[u] = p < P ? (int)partial_n[(size_t)p * Q + q] : 0;
                nmax = n[u] > nmax ? n[u] : nmax;
            }
            nmax = __reduce_max_sync(0xffffffffu, nmax);
            for (int e = 0; e < nmax; ++e) {
                unsigned long long key[MU];
#pragma unroll
                for (int u = 0; u < MU; ++u) {
                    const int p = p0 + u * 32 + lane;
                    key[u] = e < n[u] ? __ldcs(partial + ((size_t)p * Q + q) * k + e) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < MU; ++u) {
                    unsigned m = __ballot_sync(0xffffffffu, key[u] > t);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const unsigned long long x = __shfl_sync(0xffffffffu, key[u], src);
                        if (x > t) { fin.insert(x, lane); t = fin.kth(k); }
                    }
                }
            }
        }
    } else {
        const int total = P * k;
        for (int i0 = 0; i0 < total; i0 += 32 * MU) {
            unsigned long long key[MU];
#pragma unroll
            for (int u = 0; u < MU; ++u) {
                const int i = i0 + u * 32 + lane;
                key[u] = 0ull;
                if (i < total) {
                    const int p = i / k, j = i - p * k;
                    key[u] = __ldcs(partial + ((size_t)p * Q + q) * k + j);
                }
            }
#pragma unroll
            for (int u = 0; u < MU; ++u) {
                unsigned m = __ballot_sync(0xffffffffu, key[u] > t);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long x = __shfl_sync(0xffffffffu, key[u], src);
                    if (x > t) { fin.insert(x, lane); t = fin.kth(k); }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KR; ++r) {
        const int j = r * 32 + lane;
        if (j < k) {
            const unsigned long long key = fin.e[r];
            if (ko.keys != nullptr) {
                // packed candidate in global ids (score bits unchanged): 8 bytes per entry for the shard exchange
                const int64_t row = (int64_t)key64_idx(key);
                const int64_t gid = row < ko.n0 ? ko.lo0 + row : ko.lo1 + (row - ko.n0);
                ko.keys[q * k + j] = key ? ((key & 0xFFFFFFFF00000000ull) | (unsigned long long)(~(uint32_t)gid)) : 0ull;
            } else {
                out_scores[q * k + j] = key ? key64_score(key) : -INFINITY;
                out_idx[q * k + j] = key ? (int64_t)key64_idx(key) + item_id_offset : -1;
            }
        }
    }
}

// shard merge: candidates are (fp32 score, int64 global id) — [G, Q, k]
template <int KR>
__global__ void __launch_bounds__(256)
merge_cands_kernel(const float* __restrict__ cs, const int64_t* __restrict__ ci, int G, int64_t Q, int k,
                   float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    WarpList<KR> fin;
    fin.clear();
    unsigned long long t = 0ull;
    const int64_t total = (int64_t)G * k;
    for (int64_t i0 = 0; i0 < total; i0 += 32) {
        const int64_t i = i0 + lane;
        unsigned long long key = 0ull;
        if (i < total) {
            const int64_t g = i / k, j = i - g * k;
            const int64_t id = ci[((size_t)g * Q + q) * k + j];
            if (id >= 0) key = make_key64(cs[((size_t)g * Q + q) * k + j], (uint32_t)id);
        }
        unsigned m = __ballot_sync(0xffffffffu, key > t);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long x = __shfl_sync(0xffffffffu, key, src);
            if (x > t) { fin.insert(x, lane); t = fin.kth(k); }
        }
    }
#pragma unroll
    for (int r = 0; r < KR; ++r) {
        const int j = r * 32 + lane;
        if (j < k) {
            const unsigned long long key = fin.e[r];
            out_scores[q * k + j] = key ? key64_score(key) : -INFINITY;
            out_idx[q * k + j] = key ? (int64_t)key64_idx(key) : -1;
        }
    }
}

__global__ void topk_hits_kernel(const int64_t* __restrict__ topk_idx, int64_t Q, int k,
                                 const int32_t* __restrict__ pos_rowptr, const int32_t* __restrict__ pos_cols,
                                 int32_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * (k + 1)) return;
    const int64_t q = t / (k + 1);
    const int j = (int)(t - q * (k + 1));
    const int lo = pos_rowptr[q], hi = pos_rowptr[q + 1];
    if (j == k) { out[t] = hi - lo; return; }                  // pos_len column (collector.py:163)
    out[t] = sorted_contains(pos_cols, lo, hi, topk_idx[q * k + j]) ? 1 : 0;
}

// The seven collectors of inductive/evaluator.py:29-56 in one launch, one thread per (collector, user, column):
//   c: 0 overall, 1 old_users, 2 new_users, 3 old_old, 4 old_new, 5 new_old, 6 new_new   (user filter x item filter)
// out[c][q] = [hit flags of the collector's k-list | pos_len | keep].  keep = the row belongs to the collector's
// 'rec.topk' (filtered_collector.py:32-62 keeps the users that pass the user filter and still own a positive).
// The k-lists are the all / old-items-only / new-items-only top-k of ONE scoring pass (items are split at n_old_items).
// compat = 1 reproduces two quirks of collector_filter.py: :172-175 picks the blanked item segment from
// return_old_USERS instead of return_old_items, and :249-250 shifts new-item positives by -n_old_items while the
// score columns stay global.
__global__ void topk_hits_collectors_kernel(const int64_t* __restrict__ idx_all, const int64_t* __restrict__ idx_old,
                                            const int64_t* __restrict__ idx_new, int64_t Q, int k,
                                            const int64_t* __restrict__ user_ids, int64_t n_old_users, int64_t n_old_items,
                                            const int32_t* __restrict__ pos_rowptr, const int32_t* __restrict__ pos_cols,
                                            int compat, int32_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = k + 2;
    if (t >= 7 * Q * w) return;
    const int c = (int)(t / (Q * w));
    const int64_t q = (t - (int64_t)c * Q * w) / w;
    const int j = (int)(t - ((int64_t)c * Q + q) * w);
    // user filter: -1 none, 1 old users, 0 new users; item filter likewise
    const int ru = c == 0 ? -1 : (c == 1 || c == 3 || c == 4 ? 1 : 0);
    const int ri = c <= 2 ? -1 : (c == 3 || c == 5 ? 1 : 0);
    const int lo = pos_rowptr[q], hi = pos_rowptr[q + 1];
    // the positives are ascending: [lo, mid) are old items, [mid, hi) new ones
    int a = lo, b = hi;
    while (a < b) { const int m = (a + b) >> 1; if ((int64_t)pos_cols[m] < n_old_items) a = m + 1; else b = m; }
    const int mid = a;
    const int plo = ri == 0 ? mid : lo, phi = ri == 1 ? mid : hi;            // considered positives
    if (j == k) { out[t] = phi - plo; return; }                               // pos_len (collector.py:163)
    if (j == k + 1) {
        const bool old_user = user_ids[q] < n_old_users;
        const bool user_ok = ru < 0 || (ru == 1) == old_user;
        out[t] = (user_ok && (c == 0 || phi > plo)) ? 1 : 0;
        return;
    }
    const int keep_old = compat ? ru : ri;                                    // which k-list the collector reads
    const int64_t* list = ri < 0 ? idx_all : (keep_old == 1 ? idx_old : idx_new);
    int64_t id = list[q * k + j];
    if (id >= 0 && compat && ri == 0) id += n_old_items;                      // (positives shifted by -n_old_items)
    out[t] = (id >= 0 && id <= 0x7fffffffll && sorted_contains(pos_cols, plo, phi, id)) ? 1 : 0;
}

// dense scores (the reference's materialised matrix) — one warp per 32 items x TQ users
template <typename T, int TQ>
__global__ void __launch_bounds__(FS_THREADS)
fullsort_scores_simt(const T* __restrict__ users, const T* __restrict__ items, int64_t Q, int64_t N, int D,
                     int64_t item_id_offset, int mask_pad, int64_t seg_lo, int64_t seg_hi,
                     float* __restrict__ scores, int64_t ld) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* us = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * TQ;
    for (int i = tid; i < TQ * D; i += FS_THREADS) {
        const int u = i / D, d = i - u * D;
        us[i] = (q0 + u < Q) ? to_f32<T>(users[(q0 + u) * D + d]) : 0.f;
    }
    __syncthreads();
    const int64_t n_batches = (N + 31) / 32;
    for (int64_t b = (int64_t)blockIdx.x * FS_WARPS + warp; b < n_batches; b += (int64_t)gridDim.x * FS_WARPS) {
        const int64_t li = b * 32 + lane;
        const bool valid = li < N;
        const T* row = items + (valid ? li : 0) * D;
        float acc[TQ];
#pragma unroll
        for (int u = 0; u < TQ; ++u) acc[u] = 0.f;
        for (int d0 = 0; d0 < D; d0 += 16) {
            float x[16];
            load16<T>(row + d0, x);
#pragma unroll
            for (int u = 0; u < TQ; ++u) {
                const float4* up = reinterpret_cast<const float4*>(us + u * D + d0);
                float a = acc[u];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = up[j];
                    a = fmaf(x[4 * j], w.x, a); a = fmaf(x[4 * j + 1], w.y, a);
                    a = fmaf(x[4 * j + 2], w.z, a); a = fmaf(x[4 * j + 3], w.w, a);
                }
                acc[u] = a;
            }
        }
        const int64_t gid = li + item_id_offset;
        const bool item_masked = (mask_pad && gid == 0) || gid < seg_lo || gid >= seg_hi;
        if (valid) {
#pragma unroll
            for (int u = 0; u < TQ; ++u)
                if (q0 + u < Q) scores[(q0 + u) * ld + li] = item_masked ? -INFINITY : acc[u];
        }
    }
}
__global__ void hist_scatter_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, int64_t Q,
                                    int64_t N, int64_t item_id_offset, float* __restrict__ scores, int64_t ld) {
    const int64_t q = blockIdx.x;
    if (q >= Q) return;
    for (int h = rowptr[q] + threadIdx.x; h < rowptr[q + 1]; h += blockDim.x) {
        const int64_t li = (int64_t)cols[h] - item_id_offset;
        if (li >= 0 && li < N) scores[q * ld + li] = -INFINITY;
    }
}


// top-k of an already materialised (and masked) score matrix: one CTA per row — the dense entry point
// behind Collector.eval_batch_collect (collector.py:153-159) for callers that hold [Q, N] scores.
template <int KR>
__global__ void __launch_bounds__(FS_THREADS)
dense_topk_kernel(const float* __restrict__ scores, int64_t ld, int64_t N, int k,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    __shared__ unsigned long long lists[FS_WARPS * KR * 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const float* row = scores + q * ld;
    WarpList<KR> wl;
    wl.clear();
    unsigned long long t = 0ull;
    for (int64_t i0 = (int64_t)warp * 32; i0 < N; i0 += FS_THREADS) {
        const int64_t i = i0 + lane;
        const unsigned long long key = i < N ? make_key64(__ldg(row + i), (uint32_t)i) : 0ull;
        unsigned m = __ballot_sync(0xffffffffu, key > t);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long x = __shfl_sync(0xffffffffu, key, src);
            if (x > t) { wl.insert(x, lane); t = wl.kth(k); }
        }
    }
#pragma unroll
    for (int r = 0; r < KR; ++r) lists[(warp * KR + r) * 32 + lane] = wl.e[r];
    __syncthreads();
    if (warp != 0) return;
    WarpList<KR> fin;
    fin.clear();
    t = 0ull;
    for (int w = 0; w < FS_WARPS; ++w) {
#pragma unroll
        for (int r = 0; r < KR; ++r) {
            const unsigned long long key = lists[(w * KR + r) * 32 + lane];
            unsigned m = __ballot_sync(0xffffffffu, key > t);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long x = __shfl_sync(0xffffffffu, key, src);
                if (x > t) { fin.insert(x, lane); t = fin.kth(k); }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KR; ++r) {
        const int j = r * 32 + lane;
        if (j < k) {
            const unsigned long long key = fin.e[r];
            out_scores[q * k + j] = key ? key64_score(key) : -INFINITY;
            out_idx[q * k + j] = key ? (int64_t)key64_idx(key) : -1;
        }
    }
}

static int pick_item_ctas(int64_t N, int64_t user_tiles) {
    int64_t want = cdiv(N, 32 * FS_WARPS);                 // one batch per warp at least
    int64_t cap = (int64_t)num_sms() * 2;
    if (user_tiles >= 4) cap = num_sms();                  // keep the partial-list volume bounded
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

struct TopkCfg { int TQ, KR; };
static TopkCfg topk_cfg(int k) {
    if (k <= 32) return {16, 1};
    if (k <= 64) return {8, 2};
    return {4, 4};
}

template <typename T>
static int launch_fullsort_simt(const T* users, const T* items, int64_t Q, int64_t N, int D, int k,
                                int64_t off, int mask_pad, int64_t seg_lo, int64_t seg_hi,
                                const int32_t* hr, const int32_t* hc, float* out_scores, int64_t* out_idx,
                                unsigned long long* partial, int P, cudaStream_t st, const KeyOut ko = KeyOut{}) {
    const TopkCfg c = topk_cfg(k);
    const dim3 grid((unsigned)P, (unsigned)cdiv(Q, c.TQ));
    const size_t smem = (size_t)c.TQ * D * 4 + (size_t)FS_WARPS * c.TQ * c.KR * 32 * 8;
#define OOV_FS_CASE(TQ_, KR_)                                                                                         \
    {                                                                                                                 \
        auto kern = fullsort_topk_simt<T, TQ_, KR_>;                                                                  \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        kern<<<grid, FS_THREADS, smem, st>>>(users, items, Q, N, D, k, off, mask_pad, seg_lo, seg_hi, hr, hc, partial); \
        OOV_LAUNCH_CHECK("fullsort_topk_simt");                                                                       \
        merge_keys_kernel<KR_><<<(unsigned)cdiv(Q * 32, 256), 256, 0, st>>>(partial, nullptr, P, Q, k, off, out_scores, out_idx, ko); \
        OOV_LAUNCH_CHECK("merge_keys_kernel");                                                                        \
    }
    if (c.KR == 1) OOV_FS_CASE(16, 1)
    else if (c.KR == 2) OOV_FS_CASE(8, 2)
    else OOV_FS_CASE(4, 4)
#undef OOV_FS_CASE
    return OOV_OK;
}

// (row, item) pairs -> CSR with ascending columns per row, in ONE launch and without cross-CTA communication: every CTA
// owns a contiguous block of rows and reads ALL pairs twice (they stay in L2): pass 1 counts its rows' entries and the
// valid pairs that sort before its block (its rowptr base), pass 2 collects its rows' columns in shared memory, then one
// warp per row rank-sorts and writes the row out.  Replaces a torch.sort over 64-bit keys (several radix passes) on the
// per-batch path; rows outside [0, Q) are padding.
constexpr int CSR_THREADS = 1024;
constexpr int CSR_MAX_ROWS = 512;               // rows per CTA
constexpr int CSR_STAGE = 10240;                // staged columns per CTA (40 KB); more take the in-place global path

__device__ __forceinline__ void csr_sort_row(int* buf, int m, int* dst, int lane) {
    // rank sort of buf[0, m) into dst[0, m) (buf and dst may alias only if m <= 128: values are read first)
    if (m <= 128) {
        int v[4], rk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[u] = u * 32 + lane < m ? buf[u * 32 + lane] : 0x7fffffff; rk[u] = 0; }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (w * 32 >= m) break;                                       // warp-uniform
            const int jn = m - w * 32 < 32 ? m - w * 32 : 32;
            for (int jj = 0; jj < jn; ++jj) {
                const int o = __shfl_sync(0xffffffffu, v[w], jj);
                const int j = w * 32 + jj;
#pragma unroll
                for (int u = 0; u < 4; ++u) rk[u] += (o < v[u] || (o == v[u] && j < u * 32 + lane)) ? 1 : 0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (u * 32 + lane < m) dst[rk[u]] = v[u];
    } else {
        // very long rows (more than 128 history items of one user in the batch): odd-even transposition in place
        for (int pass = 0; pass < m; ++pass) {
            for (int i = (pass & 1) + 2 * lane; i + 1 < m; i += 64) {
                const int a = buf[i], c = buf[i + 1];
                if (a > c) { buf[i] = c; buf[i + 1] = a; }
            }
            __syncwarp();
        }
        if (dst != buf)
            for (int i = lane; i < m; i += 32) dst[i] = buf[i];
    }
}

// optional column map of a row shard made of two id ranges: [lo0, hi0) -> local rows 0.., [lo1, hi1) -> local rows
// (hi0 - lo0)..; every other column is dropped.  lo0 = hi0 = lo1 = 0, hi1 = INT64_MAX is the identity.
struct CsrColMap { int64_t lo0, hi0, lo1, hi1; };
__device__ __forceinline__ int64_t csr_map_col(const CsrColMap& m, int64_t c) {
    if (c >= m.lo0 && c < m.hi0) return c - m.lo0;
    if (c >= m.lo1 && c < m.hi1) return c - m.lo1 + (m.hi0 - m.lo0);
    return -1;
}

__global__ void __launch_bounds__(CSR_THREADS)
pairs_to_csr_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, int64_t n, int Q, int rows_per,
                    const CsrColMap cmap, int32_t* __restrict__ rowptr, int32_t* __restrict__ cols_out) {
    __shared__ int cnt[CSR_MAX_ROWS + 1];       // counts -> exclusive offsets inside this CTA's block of rows
    __shared__ int cur[CSR_MAX_ROWS];           // scatter cursors
    __shared__ int warp_red[CSR_THREADS / 32];
    __shared__ int stage[CSR_STAGE];
    __shared__ int s_base, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * rows_per, r1 = r0 + rows_per < Q ? r0 + rows_per : Q, nr = r1 - r0;
    for (int i = tid; i <= nr; i += CSR_THREADS) cnt[i] = 0;
    __syncthreads();
    int below = 0;
    for (int64_t i0 = 0; i0 < n; i0 += CSR_THREADS * 8) {                // 8 loads in flight per thread
        int64_t r[8], c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = i0 + u * CSR_THREADS + tid;
            r[u] = i < n ? __ldg(rows + i) : -1;
            c[u] = i < n ? __ldg(cols + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r[u] < 0 || r[u] >= Q || csr_map_col(cmap, c[u]) < 0) continue;
            if (r[u] < r0) ++below;
            else if (r[u] < r1) atomicAdd(&cnt[(int)r[u] - r0], 1);
        }
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 0) warp_red[warp] = below;
    __syncthreads();
    if (warp == 0) {
        const int b = __reduce_add_sync(0xffffffffu, warp_red[lane]);
        // exclusive scan of this block's <= 512 row counts: 16 per lane
        int c[CSR_MAX_ROWS / 32], local = 0;
#pragma unroll
        for (int u = 0; u < CSR_MAX_ROWS / 32; ++u) { const int i = lane * (CSR_MAX_ROWS / 32) + u; c[u] = i < nr ? cnt[i] : 0; local += c[u]; }
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        int run = incl - local;
#pragma unroll
        for (int u = 0; u < CSR_MAX_ROWS / 32; ++u) {
            const int i = lane * (CSR_MAX_ROWS / 32) + u;
            if (i < nr) { cnt[i] = run; cur[i] = run; }
            run += c[u];
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) { s_base = b; s_total = total; cnt[nr] = total; }
    }
    __syncthreads();
    const int base = s_base, total = s_total;
    for (int i = tid; i < nr; i += CSR_THREADS) rowptr[r0 + i] = base + cnt[i];
    if (r1 == Q && tid == 0) rowptr[Q] = base + total;
    const bool staged = total <= CSR_STAGE;
    int* buf = staged ? stage : cols_out + base;
    for (int64_t i0 = 0; i0 < n && total > 0; i0 += CSR_THREADS * 8) {
        int64_t r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int64_t i = i0 + u * CSR_THREADS + tid; r[u] = i < n ? __ldg(rows + i) : -1; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r[u] < r0 || r[u] >= r1) continue;
            const int64_t c = csr_map_col(cmap, __ldg(cols + i0 + u * CSR_THREADS + tid));
            if (c >= 0) buf[atomicAdd(&cur[(int)r[u] - r0], 1)] = (int32_t)c;
        }
    }
    __syncthreads();                            // (global path: a CTA sees its own writes after the barrier)
    for (int i = warp; i < nr; i += CSR_THREADS / 32) {
        const int lo = cnt[i], m = cnt[i + 1] - lo;
        if (m == 0) continue;
        if (staged) csr_sort_row(stage + lo, m, cols_out + base + lo, lane);
        else if (m > 1) csr_sort_row(cols_out + base + lo, m, cols_out + base + lo, lane);
    }
}

// shared with the tcgen05 scoring kernel (tc_score.cu): merge P partial key lists per user
int launch_merge_keys(const unsigned long long* partial, const uint8_t* partial_n, int P, int64_t Q, int k, int64_t off,
                      float* out_scores, int64_t* out_idx, cudaStream_t st, const KeyOut ko) {
    const unsigned blocks = (unsigned)cdiv(Q * 32, 256);
    if (k <= 32) merge_keys_kernel<1><<<blocks, 256, 0, st>>>(partial, partial_n, P, Q, k, off, out_scores, out_idx, ko);
    else if (k <= 64) merge_keys_kernel<2><<<blocks, 256, 0, st>>>(partial, partial_n, P, Q, k, off, out_scores, out_idx, ko);
    else merge_keys_kernel<4><<<blocks, 256, 0, st>>>(partial, partial_n, P, Q, k, off, out_scores, out_idx, ko);
    OOV_LAUNCH_CHECK("merge_keys_kernel");
    return OOV_OK;
}

namespace tc {
bool score_tc_supported(int dtype, int D, int k);
size_t score_tc_workspace(int64_t Q, int64_t N, int k);
int score_tc_run(const void* users, const void* items, int64_t Q, int64_t N, int D, int k, int64_t item_id_offset,
                 int mask_pad, int64_t seg_lo, int64_t seg_hi, const int32_t* hist_rowptr, const int32_t* hist_cols,
                 float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t st,
                 const KeyOut ko = KeyOut{});
}  // namespace tc

}  // namespace oov


// ------------------------------------------------------------------------------------ sampled-candidate evaluation
// Replaces trainer.py:547-564 / inductive/evaluator.py:116-133 (neg_sample_batch_eval): origin_scores = model.predict(
// (user, item) pairs) = (user_e * item_e).sum(-1) (bpr.py:146-149), scattered into a [users, N] matrix of -inf, then
// torch.topk per user.  Here the [users, N] matrix never exists: the pairs arrive as a CSR (row = batch user, cols = the
// candidate item ids ascending — oov_pairs_to_csr), item_e holds one embedding row per CSR entry, and a row's top-k is
// selected among ITS candidates only (every other item scores -inf in the reference and can only fill the tail of a row
// with fewer than k candidates: those slots come back as (-inf, -1)).  Duplicate pairs of a row collapse like the
// scatter's last-write-wins (equal ids are adjacent in the sorted CSR row).
//   pair_keys_kernel      : lane = CSR entry of the warp's row: fp32 dot product -> key64 = order_key(score) << 32 | ~id
//                           (0 for a duplicate).
//   rows_topk_keys_kernel : warp = row; k rounds of "largest key strictly below the previous winner" among ids inside
//                           [seg_lo, seg_hi) (keys are distinct within a row, so no entry needs marking) — (score desc,
//                           id asc), deterministic.
namespace oov {
__global__ void __launch_bounds__(256)
pair_keys_kernel(const void* __restrict__ user_e, int u_dtype, const void* __restrict__ item_e, int i_dtype, int D,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, int64_t U, int normalize,
                 unsigned long long* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= U) return;
    const int b = rowptr[row], e = rowptr[row + 1];
    for (int j = b + lane; j < e; j += 32) {
        const int32_t id = cols[j];
        if (j > b && cols[j - 1] == id) { keys[j] = 0ull; continue; }
        float acc = 0.f;
        if (normalize) {
            // DirectAU.predict (directau.py:75-78,174-181): F.normalize both rows (x / max(|x|_2, 1e-12)), then multiply-sum
            float su = 0.f, si = 0.f;
            for (int d = 0; d < D; ++d) {
                const float u = load_elem(user_e, u_dtype, row * D + d), v = load_elem(item_e, i_dtype, (int64_t)j * D + d);
                su = fmaf(u, u, su); si = fmaf(v, v, si);
            }
            const float nu = fmaxf(sqrtf(su), 1e-12f), ni = fmaxf(sqrtf(si), 1e-12f);
            for (int d = 0; d < D; ++d)
                acc += (load_elem(user_e, u_dtype, row * D + d) / nu) * (load_elem(item_e, i_dtype, (int64_t)j * D + d) / ni);
        } else {
            for (int d = 0; d < D; ++d)                               // the user row is a broadcast load, the item row this lane's own
                acc = fmaf(load_elem(user_e, u_dtype, row * D + d), load_elem(item_e, i_dtype, (int64_t)j * D + d), acc);
        }
        keys[j] = ((unsigned long long)float_order_key(acc) << 32) | (unsigned long long)(~(uint32_t)id);
    }
}

__global__ void __launch_bounds__(256)
rows_topk_keys_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ rowptr, int64_t U, int k,
                      int64_t seg_lo, int64_t seg_hi, float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= U) return;
    const int b = rowptr[row], e = rowptr[row + 1];
    unsigned long long last = ~0ull;
    for (int r = 0; r < k; ++r) {
        unsigned long long best = 0ull;
        if (last != 0ull) {
            for (int j = b + lane; j < e; j += 32) {
                const unsigned long long x = keys[j];
                const int64_t id = (int64_t)(~(uint32_t)x);
                if (x != 0ull && x < last && x > best && id >= seg_lo && id < seg_hi) best = x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
                best = y > best ? y : best;
            }
        }
        last = best;
        if (lane == 0) {
            out_scores[row * k + r] = best ? float_from_order_key((uint32_t)(best >> 32)) : -INFINITY;
            out_idx[row * k + r] = best ? (int64_t)(~(uint32_t)best) : -1;
        }
    }
}
}  // namespace oov

using namespace oov;

extern "C" {

size_t oov_fullsort_topk_workspace(int64_t Q, int64_t N, int32_t D, int32_t k, int32_t path) {
    if (Q <= 0 || N <= 0 || k <= 0) return 0;
    const TopkCfg c = topk_cfg(k);
    const int P = pick_item_ctas(N, cdiv(Q, c.TQ));
    size_t a = align_up((size_t)P * Q * k * 8, 256);
    if (path != OOV_PATH_SIMT_FP32 && tc::score_tc_supported(OOV_BF16, D, k)) {
        const size_t b = tc::score_tc_workspace(Q, N, k);
        if (b > a) a = b;
    }
    return a;
}

int oov_pairs_to_csr(const int64_t* rows, const int64_t* cols, int64_t n_pairs, int64_t Q, const int64_t* col_ranges,
                     int32_t* rowptr_out, int32_t* cols_out, void* stream) {
    OOV_REQUIRE(Q >= 1 && Q <= (1ll << 24) && n_pairs >= 0 && n_pairs < (1ll << 31), OOV_ERR_ARG,
                "oov_pairs_to_csr: Q=%lld (1..2^24), n_pairs=%lld (< 2^31)", (long long)Q, (long long)n_pairs);
    OOV_REQUIRE(rowptr_out && (n_pairs == 0 || (rows && cols && cols_out)), OOV_ERR_ARG, "oov_pairs_to_csr: NULL pointer");
    // 64-512 rows per CTA: every CTA streams all the pairs (L2-resident after the first), so up to 8192 rows use at most 16
    // CTAs; larger query batches take ceil(Q / 512) CTAs
    int rows_per = (int)cdiv(Q, 16);
    if (rows_per < 64) rows_per = 64;
    if (rows_per > CSR_MAX_ROWS) rows_per = CSR_MAX_ROWS;
    const unsigned grid = (unsigned)cdiv(Q, rows_per);
    CsrColMap cmap{0, 0, 0, INT64_MAX};                               // identity
    if (col_ranges != nullptr) {
        cmap = CsrColMap{col_ranges[0], col_ranges[1], col_ranges[2], col_ranges[3]};
        OOV_REQUIRE(cmap.lo0 <= cmap.hi0 && cmap.hi0 <= cmap.lo1 && cmap.lo1 <= cmap.hi1, OOV_ERR_ARG,
                    "oov_pairs_to_csr: col_ranges must be two ascending, disjoint ranges");
    }
    pairs_to_csr_kernel<<<grid, CSR_THREADS, 0, (cudaStream_t)stream>>>(rows, cols, n_pairs, (int)Q, rows_per, cmap, rowptr_out, cols_out);
    OOV_LAUNCH_CHECK("pairs_to_csr_kernel");
    return OOV_OK;
}

static int fullsort_topk_impl(const void* users, const void* items, int32_t dtype, int64_t Q, int64_t N, int32_t D, int32_t k,
                              int64_t item_id_offset, int32_t mask_pad, int64_t seg_lo, int64_t seg_hi,
                              const int32_t* hist_rowptr, const int32_t* hist_cols, float* out_scores, int64_t* out_idx,
                              void* workspace, size_t workspace_bytes, int32_t path, void* stream, const KeyOut ko) {
    OOV_REQUIRE(dtype_ok(dtype), OOV_ERR_ARG, "oov_fullsort_topk: bad dtype %d", dtype);
    OOV_REQUIRE(Q >= 0 && N >= 0 && D > 0 && k > 0 && k <= 128, OOV_ERR_ARG,
                "oov_fullsort_topk: bad shape Q=%lld N=%lld D=%d k=%d (k <= 128)", (long long)Q, (long long)N, D, k);
    OOV_REQUIRE(N < (1ll << 32), OOV_ERR_ARG, "oov_fullsort_topk: shard has %lld rows (max 2^32-1)", (long long)N);
    OOV_REQUIRE(D % 16 == 0, OOV_ERR_ARG, "oov_fullsort_topk: D=%d must be a multiple of 16 (pad the tables)", D);
    OOV_REQUIRE((hist_rowptr == nullptr) == (hist_cols == nullptr), OOV_ERR_ARG, "oov_fullsort_topk: rowptr/cols mismatch");
    OOV_REQUIRE(path >= OOV_PATH_AUTO && path <= OOV_PATH_TCGEN05, OOV_ERR_ARG, "oov_fullsort_topk: unsupported path %d", path);
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(users && ((out_scores && out_idx) || ko.keys) && (N == 0 || items), OOV_ERR_ARG, "oov_fullsort_topk: NULL pointer");
    OOV_REQUIRE(aligned(users, 16) && aligned(items, 16), OOV_ERR_ALIGN, "oov_fullsort_topk: tables must be 16-byte aligned");
    // bf16 tables take the tensor-core path (tcgen05 GEMM fused with the top-k epilogue); fp32 tables the fp32 FMA path
    const bool tc_ok = tc::score_tc_supported(dtype, D, k) && N >= 1;
    OOV_REQUIRE(path != OOV_PATH_TCGEN05 || tc_ok, OOV_ERR_ARG,
                "oov_fullsort_topk: tcgen05 path needs bf16 tables, D <= 64 (multiple of 8) and k <= 24");
    if (tc_ok && path != OOV_PATH_SIMT_FP32)
        return tc::score_tc_run(users, items, Q, N, D, k, item_id_offset, mask_pad, seg_lo, seg_hi, hist_rowptr, hist_cols,
                                out_scores, out_idx, workspace, workspace_bytes, (cudaStream_t)stream, ko);
    const TopkCfg c = topk_cfg(k);
    const int P = pick_item_ctas(N > 0 ? N : 1, cdiv(Q, c.TQ));
    const size_t need = (size_t)P * Q * k * 8;
    OOV_REQUIRE(workspace && workspace_bytes >= need, OOV_ERR_WORKSPACE, "oov_fullsort_topk: workspace %zu < %zu",
                workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* partial = reinterpret_cast<unsigned long long*>(workspace);
    if (dtype == OOV_F32)
        return launch_fullsort_simt<float>((const float*)users, (const float*)items, Q, N, D, k, item_id_offset, mask_pad,
                                           seg_lo, seg_hi, hist_rowptr, hist_cols, out_scores, out_idx, partial, P, st, ko);
    return launch_fullsort_simt<__nv_bfloat16>((const __nv_bfloat16*)users, (const __nv_bfloat16*)items, Q, N, D, k,
                                               item_id_offset, mask_pad, seg_lo, seg_hi, hist_rowptr, hist_cols,
                                               out_scores, out_idx, partial, P, st, ko);
}

int oov_fullsort_topk(const void* users, const void* items, int32_t dtype, int64_t Q, int64_t N, int32_t D, int32_t k,
                      int64_t item_id_offset, int32_t mask_pad, int64_t seg_lo, int64_t seg_hi,
                      const int32_t* hist_rowptr, const int32_t* hist_cols, float* out_scores, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    return fullsort_topk_impl(users, items, dtype, Q, N, D, k, item_id_offset, mask_pad, seg_lo, seg_hi, hist_rowptr, hist_cols,
                              out_scores, out_idx, workspace, workspace_bytes, path, stream, KeyOut{});
}

int oov_fullsort_topk_keys(const void* users, const void* items, int32_t dtype, int64_t Q, int64_t N, int32_t D, int32_t k,
                           int32_t mask_pad, int64_t seg_lo, int64_t seg_hi, const int32_t* hist_rowptr, const int32_t* hist_cols,
                           int64_t n0, int64_t lo0, int64_t lo1, uint64_t* keys_out,
                           void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    OOV_REQUIRE(keys_out != nullptr || Q == 0, OOV_ERR_ARG, "oov_fullsort_topk_keys: keys_out is NULL");
    OOV_REQUIRE(n0 >= 0 && n0 <= N && lo0 >= 0 && lo1 >= 0 && lo0 + n0 < (1ll << 32) && lo1 + (N - n0) < (1ll << 32), OOV_ERR_ARG,
                "oov_fullsort_topk_keys: row map n0=%lld lo0=%lld lo1=%lld does not fit 32-bit global ids", (long long)n0,
                (long long)lo0, (long long)lo1);
    KeyOut ko;
    ko.keys = reinterpret_cast<unsigned long long*>(keys_out); ko.n0 = n0; ko.lo0 = lo0; ko.lo1 = lo1;
    return fullsort_topk_impl(users, items, dtype, Q, N, D, k, 0, mask_pad, seg_lo, seg_hi, hist_rowptr, hist_cols, nullptr, nullptr,
                              workspace, workspace_bytes, path, stream, ko);
}

int oov_topk_merge_keys(const uint64_t* keys, int32_t G, int64_t Q, int32_t k, float* out_scores, int64_t* out_idx, void* stream) {
    OOV_REQUIRE(G > 0 && Q >= 0 && k > 0 && k <= 128, OOV_ERR_ARG, "oov_topk_merge_keys: bad shape G=%d Q=%lld k=%d", G, (long long)Q, k);
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(keys && out_scores && out_idx, OOV_ERR_ARG, "oov_topk_merge_keys: NULL pointer");
    return launch_merge_keys(reinterpret_cast<const unsigned long long*>(keys), nullptr, G, Q, k, 0, out_scores, out_idx,
                             (cudaStream_t)stream, KeyOut{});
}

int oov_fullsort_scores(const void* users, const void* items, int32_t dtype, int64_t Q, int64_t N, int32_t D,
                        int64_t item_id_offset, int32_t mask_pad, int64_t seg_lo, int64_t seg_hi,
                        const int32_t* hist_rowptr, const int32_t* hist_cols, float* scores, int64_t scores_stride,
                        void* stream) {
    OOV_REQUIRE(dtype_ok(dtype) && Q >= 0 && N >= 0 && D > 0 && D % 16 == 0 && scores_stride >= N, OOV_ERR_ARG,
                "oov_fullsort_scores: bad argument (D must be a multiple of 16)");
    OOV_REQUIRE((hist_rowptr == nullptr) == (hist_cols == nullptr), OOV_ERR_ARG, "oov_fullsort_scores: rowptr/cols mismatch");
    if (Q == 0 || N == 0) return OOV_OK;
    OOV_REQUIRE(users && items && scores, OOV_ERR_ARG, "oov_fullsort_scores: NULL pointer");
    OOV_REQUIRE(aligned(users, 16) && aligned(items, 16), OOV_ERR_ALIGN, "oov_fullsort_scores: tables must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int TQ = 16;
    const dim3 grid((unsigned)pick_item_ctas(N, 1), (unsigned)cdiv(Q, TQ));
    const size_t smem = (size_t)TQ * D * 4;
    if (dtype == OOV_F32) {
        auto kern = fullsort_scores_simt<float, TQ>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, FS_THREADS, smem, st>>>((const float*)users, (const float*)items, Q, N, D, item_id_offset, mask_pad,
                                             seg_lo, seg_hi, scores, scores_stride);
    } else {
        auto kern = fullsort_scores_simt<__nv_bfloat16, TQ>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, FS_THREADS, smem, st>>>((const __nv_bfloat16*)users, (const __nv_bfloat16*)items, Q, N, D,
                                             item_id_offset, mask_pad, seg_lo, seg_hi, scores, scores_stride);
    }
    OOV_LAUNCH_CHECK("fullsort_scores_simt");
    if (hist_rowptr != nullptr) {
        hist_scatter_kernel<<<(unsigned)Q, 64, 0, st>>>(hist_rowptr, hist_cols, Q, N, item_id_offset, scores, scores_stride);
        OOV_LAUNCH_CHECK("hist_scatter_kernel");
    }
    return OOV_OK;
}

int oov_dense_topk(const float* scores, int64_t scores_stride, int64_t Q, int64_t N, int32_t k,
                   float* out_scores, int64_t* out_idx, void* stream) {
    OOV_REQUIRE(Q >= 0 && N >= 0 && k > 0 && k <= 128 && scores_stride >= N && N < (1ll << 32), OOV_ERR_ARG,
                "oov_dense_topk: bad shape Q=%lld N=%lld k=%d", (long long)Q, (long long)N, k);
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(out_scores && out_idx && (N == 0 || scores), OOV_ERR_ARG, "oov_dense_topk: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 32) dense_topk_kernel<1><<<(unsigned)Q, FS_THREADS, 0, st>>>(scores, scores_stride, N, k, out_scores, out_idx);
    else if (k <= 64) dense_topk_kernel<2><<<(unsigned)Q, FS_THREADS, 0, st>>>(scores, scores_stride, N, k, out_scores, out_idx);
    else dense_topk_kernel<4><<<(unsigned)Q, FS_THREADS, 0, st>>>(scores, scores_stride, N, k, out_scores, out_idx);
    OOV_LAUNCH_CHECK("dense_topk_kernel");
    return OOV_OK;
}

int oov_topk_merge(const float* cand_scores, const int64_t* cand_idx, int32_t G, int64_t Q, int32_t k,
                   float* out_scores, int64_t* out_idx, void* stream) {
    OOV_REQUIRE(G > 0 && Q >= 0 && k > 0 && k <= 128, OOV_ERR_ARG, "oov_topk_merge: bad shape");
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(cand_scores && cand_idx && out_scores && out_idx, OOV_ERR_ARG, "oov_topk_merge: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)cdiv(Q * 32, 256);
    if (k <= 32) merge_cands_kernel<1><<<blocks, 256, 0, st>>>(cand_scores, cand_idx, G, Q, k, out_scores, out_idx);
    else if (k <= 64) merge_cands_kernel<2><<<blocks, 256, 0, st>>>(cand_scores, cand_idx, G, Q, k, out_scores, out_idx);
    else merge_cands_kernel<4><<<blocks, 256, 0, st>>>(cand_scores, cand_idx, G, Q, k, out_scores, out_idx);
    OOV_LAUNCH_CHECK("merge_cands_kernel");
    return OOV_OK;
}

int oov_topk_hits(const int64_t* topk_idx, int64_t Q, int32_t k, const int32_t* pos_rowptr, const int32_t* pos_cols,
                  int32_t* out_hits, void* stream) {
    OOV_REQUIRE(Q >= 0 && k > 0, OOV_ERR_ARG, "oov_topk_hits: bad shape");
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(topk_idx && pos_rowptr && pos_cols && out_hits, OOV_ERR_ARG, "oov_topk_hits: NULL pointer");
    topk_hits_kernel<<<(unsigned)cdiv(Q * (k + 1), 256), 256, 0, (cudaStream_t)stream>>>(topk_idx, Q, k, pos_rowptr,
                                                                                        pos_cols, out_hits);
    OOV_LAUNCH_CHECK("topk_hits_kernel");
    return OOV_OK;
}

int oov_topk_hits_collectors(const int64_t* idx_all, const int64_t* idx_old, const int64_t* idx_new, int64_t Q, int32_t k,
                             const int64_t* user_ids, int64_t n_old_users, int64_t n_old_items, const int32_t* pos_rowptr,
                             const int32_t* pos_cols, int32_t reference_compat, int32_t* out, void* stream) {
    OOV_REQUIRE(Q >= 0 && k > 0, OOV_ERR_ARG, "oov_topk_hits_collectors: bad shape");
    if (Q == 0) return OOV_OK;
    OOV_REQUIRE(idx_all && idx_old && idx_new && user_ids && pos_rowptr && pos_cols && out, OOV_ERR_ARG,
                "oov_topk_hits_collectors: NULL pointer");
    topk_hits_collectors_kernel<<<(unsigned)cdiv(7 * Q * (k + 2), 256), 256, 0, (cudaStream_t)stream>>>(
        idx_all, idx_old, idx_new, Q, k, user_ids, n_old_users, n_old_items, pos_rowptr, pos_cols, reference_compat ? 1 : 0, out);
    OOV_LAUNCH_CHECK("topk_hits_collectors_kernel");
    return OOV_OK;
}

int oov_pair_topk(const void* user_e, int32_t u_dtype, const void* item_e, int32_t i_dtype, int32_t D,
                  const int32_t* rowptr, const int32_t* cols, int64_t U, int64_t n_pairs, int32_t normalize, int32_t k, int64_t seg_lo, int64_t seg_hi,
                  unsigned long long* keys, int32_t compute_keys, float* out_scores, int64_t* out_idx, void* stream) {
    OOV_REQUIRE(U >= 0 && n_pairs >= 0 && n_pairs < (1ll << 31) && D > 0 && k > 0 && dtype_ok(u_dtype) && dtype_ok(i_dtype), OOV_ERR_ARG,
                "oov_pair_topk: bad argument");
    if (U == 0) return OOV_OK;
    OOV_REQUIRE(rowptr && out_scores && out_idx && (n_pairs == 0 || (user_e && item_e && cols && keys)), OOV_ERR_ARG, "oov_pair_topk: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (compute_keys && n_pairs > 0) {
        pair_keys_kernel<<<(unsigned)cdiv(U, 8), 256, 0, st>>>(user_e, u_dtype, item_e, i_dtype, D, rowptr, cols, U, normalize ? 1 : 0, keys);
        OOV_LAUNCH_CHECK("pair_keys_kernel");
    }
    rows_topk_keys_kernel<<<(unsigned)cdiv(U, 8), 256, 0, st>>>(keys, rowptr, U, k, seg_lo, seg_hi, out_scores, out_idx);
    OOV_LAUNCH_CHECK("rows_topk_keys_kernel");
    return OOV_OK;
}

}  // extern "C"
