// LSH / SLSH OOV embedders — CUDA-core (fp32 FMA, exact-sign) path.
//
//   lsh_bits_simt   : R = X[ids] @ P^T, bit = !(R < 0), warp-ballot packed into uint32 words
//   lsh_mean_simt   : out = (H @ W) / popcount(H) from the packed words, W staged in smem
//   slsh_embed_simt : <= 32 planes, one warp per row: ballot -> popcount -> bucket -> row gather
//
// Semantics restated from inductive/torch_hash.py:55-60, lsh_embedder.py:116-179,
// single_lsh_embedder.py:77-109 of the reference.  The tensor-core (tcgen05) LSH path lives in
// tc_lsh.cu; this file is the fp32-exact path and the path for shapes the MMA tiles do not cover.
#include "common.cuh"

namespace oov {

constexpr int BITS_TM = 32;        // rows per CTA in lsh_bits_simt
constexpr int BITS_THREADS = 256;  // one plane per thread per chunk
constexpr int BITS_FCH = 256;      // feature columns staged in smem per pass

// xs layout: [BITS_TM][BITS_FCH] floats (row r at xs + r * BITS_FCH); all lanes of a warp read the
// same address (they work on the same row at the same time) -> broadcast, conflict-free.
__global__ void __launch_bounds__(BITS_THREADS, 2)
lsh_bits_simt(const float* __restrict__ feat, int64_t n_feat_rows, int F,
              const float* __restrict__ planes, int B,
              const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n,
              int64_t n_old, int64_t prime_pad, float tie_eps,
              uint32_t* __restrict__ bits, int words, unsigned long long* __restrict__ tie_count) {
    __shared__ __align__(16) float xs[BITS_TM * BITS_FCH];
    __shared__ int64_t srow[BITS_TM];   // feature row index, or -1 for rows that need no hash
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * BITS_TM;

    if (tid < BITS_TM) {
        int64_t r = row0 + tid, fr = -1;
        if (r < n) {
            int64_t id = ids[r * ids_stride];
            if (id >= n_old) {
                fr = feature_row(id, prime_pad);
                if (fr < 0 || fr >= n_feat_rows) fr = -1;   // out-of-range ids hash nothing (caller bug)
            }
        }
        srow[tid] = fr;
    }
    __syncthreads();

    unsigned int my_ties = 0;
    const int n_pchunks = (B + BITS_THREADS - 1) / BITS_THREADS;
    for (int pc = 0; pc < n_pchunks; ++pc) {
        const int b = pc * BITS_THREADS + tid;           // this thread's plane
        const bool bvalid = b < B;
        const float* prow = planes + (size_t)(bvalid ? b : 0) * F;
        float acc[BITS_TM];
#pragma unroll
        for (int r = 0; r < BITS_TM; ++r) acc[r] = 0.f;

        for (int f0 = 0; f0 < F; f0 += BITS_FCH) {
            const int fw = min(BITS_FCH, F - f0);
            const int fw4 = (fw + 3) & ~3;
            __syncthreads();                              // previous xs consumers done
            for (int i = tid; i < BITS_TM * fw4; i += BITS_THREADS) {
                const int r = i / fw4, f = i - r * fw4;
                float v = 0.f;
                const int64_t fr = srow[r];
                if (fr >= 0 && f < fw) v = __ldg(feat + fr * F + f0 + f);
                xs[r * BITS_FCH + f] = v;
            }
            __syncthreads();
            for (int s0 = 0; s0 < fw4; s0 += 32) {
                float p[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int f = f0 + s0 + j;
                    p[j] = (bvalid && s0 + j < fw) ? __ldg(prow + f) : 0.f;
                }
                const int sw = min(32, fw4 - s0);        // multiple of 4
#pragma unroll
                for (int r = 0; r < BITS_TM; ++r) {
                    const float4* xr = reinterpret_cast<const float4*>(xs + r * BITS_FCH + s0);
                    float a = acc[r];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        if (j4 * 4 < sw) {
                            const float4 x = xr[j4];
                            a = fmaf(x.x, p[j4 * 4 + 0], a);
                            a = fmaf(x.y, p[j4 * 4 + 1], a);
                            a = fmaf(x.z, p[j4 * 4 + 2], a);
                            a = fmaf(x.w, p[j4 * 4 + 3], a);
                        }
                    }
                    acc[r] = a;
                }
            }
        }
        // torch_hash.py:57-59: R < 0 -> 0, everything else (incl. -0, NaN) -> 1
#pragma unroll
        for (int r = 0; r < BITS_TM; ++r) {
            const bool live = bvalid && srow[r] >= 0;
            const unsigned w = __ballot_sync(0xffffffffu, live && !(acc[r] < 0.f));
            if (live && fabsf(acc[r]) < tie_eps) ++my_ties;
            const int widx = pc * (BITS_THREADS / 32) + warp;
            if (lane == 0 && row0 + r < n && widx < words) bits[(row0 + r) * words + widx] = w;
        }
    }
    if (tie_count != nullptr) {
        for (int o = 16; o; o >>= 1) my_ties += __shfl_xor_sync(0xffffffffu, my_ties, o);
        if (lane == 0 && my_ties) atomicAdd(tie_count, (unsigned long long)my_ties);
    }
}

// ------------------------------------------------------------------------------------
constexpr int MEAN_TM = 64;       // rows per CTA
constexpr int MEAN_TB = 128;      // buckets staged per pass
constexpr int MEAN_TD = 64;       // embedding columns per pass
constexpr int MEAN_THREADS = 256;

__global__ void __launch_bounds__(MEAN_THREADS, 2)
lsh_mean_simt(const uint32_t* __restrict__ bits, int words, int B,
              const void* __restrict__ W, int w_dtype, int D,
              const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n, int64_t n_old,
              const void* __restrict__ iv_table, int iv_dtype,
              void* __restrict__ out, int out_dtype, int64_t out_stride) {
    __shared__ __align__(16) float Ws[MEAN_TB * MEAN_TD];       // 32 KB
    __shared__ uint32_t hs[MEAN_TM * (MEAN_TB / 32)];           // 1 KB
    __shared__ int64_t sid[MEAN_TM];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;                     // 16 row-groups x 16 column-groups
    const int64_t row0 = (int64_t)blockIdx.x * MEAN_TM;

    if (tid < MEAN_TM) {
        const int64_t r = row0 + tid;
        sid[tid] = r < n ? ids[r * ids_stride] : INT64_MIN;      // INT64_MIN = no such row
    }
    __syncthreads();

    // in-vocab rows: plain gather (bpr.py:111-112), 16 lanes per row
    for (int r = ty; r < MEAN_TM; r += 16) {
        const int64_t id = sid[r];
        if (id != INT64_MIN && id < n_old && iv_table != nullptr && id >= 0)
            copy_row(reinterpret_cast<const char*>(iv_table) + (size_t)id * D * (iv_dtype == OOV_F32 ? 4 : 2), iv_dtype,
                     reinterpret_cast<char*>(out) + (size_t)(row0 + r) * out_stride * (out_dtype == OOV_F32 ? 4 : 2),
                     out_dtype, D, tx, 16);
    }

    bool any_oov = false;
    bool mine_oov[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t id = sid[ty * 4 + i];
        mine_oov[i] = (id != INT64_MIN) && id >= n_old;
    }
    for (int r = 0; r < MEAN_TM; ++r) any_oov |= (sid[r] != INT64_MIN && sid[r] >= n_old);
    if (!any_oov) return;                                        // block-uniform

    const int n_bchunks = (B + MEAN_TB - 1) / MEAN_TB;
    for (int d0 = 0; d0 < D; d0 += MEAN_TD) {
        float tot[4][4];
        int cnt[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[i][j] = 0.f;

        for (int bc = 0; bc < n_bchunks; ++bc) {
            __syncthreads();
            for (int i = tid; i < MEAN_TB * MEAN_TD; i += MEAN_THREADS) {
                const int bb = i / MEAN_TD, dd = i - bb * MEAN_TD;
                const int b = bc * MEAN_TB + bb, d = d0 + dd;
                Ws[i] = (b < B && d < D) ? load_elem(W, w_dtype, (int64_t)b * D + d) : 0.f;
            }
            for (int i = tid; i < MEAN_TM * (MEAN_TB / 32); i += MEAN_THREADS) {
                const int r = i / (MEAN_TB / 32), w = i - r * (MEAN_TB / 32);
                const int widx = bc * (MEAN_TB / 32) + w;
                const int64_t id = sid[r];
                hs[i] = (id != INT64_MIN && id >= n_old && widx < words) ? bits[(row0 + r) * words + widx] : 0u;
            }
            __syncthreads();
            float part[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
            for (int wg = 0; wg < MEAN_TB / 32; ++wg) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[i] = hs[(ty * 4 + i) * (MEAN_TB / 32) + wg];
                    cnt[i] += __popc(w[i]);
                }
                if ((w[0] | w[1] | w[2] | w[3]) == 0u) continue;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const float4 wv = *reinterpret_cast<const float4*>(Ws + (wg * 32 + j) * MEAN_TD + tx * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((w[i] >> j) & 1u) {
                            part[i][0] += wv.x; part[i][1] += wv.y; part[i][2] += wv.z; part[i][3] += wv.w;
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) tot[i][j] += part[i][j];
        }
        // (H @ W) / H.sum(1): 0/0 -> NaN exactly like lsh_embedder.py:158
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!mine_oov[i]) continue;
            const float den = (float)cnt[i];
            const int64_t orow = (row0 + ty * 4 + i) * out_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = d0 + tx * 4 + j;
                if (d < D) store_elem(out, out_dtype, orow + d, tot[i][j] / den);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
constexpr int SLSH_THREADS = 256;
constexpr int SLSH_MAXF = 1024;   // planes^T staged whole in smem: F * 32 * 4 B <= 128 KB

__global__ void __launch_bounds__(SLSH_THREADS)
slsh_embed_simt(const float* __restrict__ feat, int64_t n_feat_rows, int F,
                const float* __restrict__ planes, int bits_req, int n_buckets,
                const void* __restrict__ W, int w_dtype,
                const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n,
                int64_t n_old, int64_t prime_pad,
                const void* __restrict__ iv_table, int iv_dtype,
                void* __restrict__ out, int out_dtype, int64_t out_stride, int D,
                float tie_eps, int64_t* __restrict__ bucket_out, unsigned long long* __restrict__ tie_count) {
    extern __shared__ float Pt[];            // [F][32], plane b of feature f at Pt[f * 32 + b]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < F * 32; i += SLSH_THREADS) {
        const int f = i >> 5, b = i & 31;
        Pt[i] = b < bits_req ? __ldg(planes + (size_t)b * F + f) : 0.f;
    }
    __syncthreads();
    const int warps_per_block = SLSH_THREADS / 32;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2;
    unsigned int my_ties = 0;
    for (int64_t r = (int64_t)blockIdx.x * warps_per_block + warp; r < n; r += (int64_t)gridDim.x * warps_per_block) {
        const int64_t id = ids[r * ids_stride];
        char* orow = out ? reinterpret_cast<char*>(out) + (size_t)r * out_stride * osz : nullptr;
        if (id < n_old) {
            if (bucket_out && lane == 0) bucket_out[r] = -1;
            if (iv_table != nullptr && orow != nullptr && id >= 0)
                copy_row(reinterpret_cast<const char*>(iv_table) + (size_t)id * D * (iv_dtype == OOV_F32 ? 4 : 2), iv_dtype,
                         orow, out_dtype, D, lane, 32);
            continue;
        }
        const int64_t fr = feature_row(id, prime_pad);
        const bool inrange = fr >= 0 && fr < n_feat_rows;
        const float* x = feat + (inrange ? fr : 0) * F;
        float acc = 0.f;
        for (int f0 = 0; f0 < F; f0 += 32) {
            const float xv = (inrange && f0 + lane < F) ? __ldg(x + f0 + lane) : 0.f;
            const int fw = min(32, F - f0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < fw) acc = fmaf(__shfl_sync(0xffffffffu, xv, j), Pt[(f0 + j) * 32 + lane], acc);
            }
        }
        const bool pl = lane < bits_req;
        const unsigned word = __ballot_sync(0xffffffffu, pl && !(acc < 0.f));
        const unsigned tie = __ballot_sync(0xffffffffu, pl && fabsf(acc) < tie_eps);
        if (lane == 0) my_ties += __popc(tie);
        // (2 ** H).sum(1) % n_buckets with H in {0,1}  ==  (bits_req + popcount) % n_buckets
        const int bucket = (bits_req + __popc(word)) % n_buckets;
        if (bucket_out && lane == 0) bucket_out[r] = bucket;
        if (W != nullptr && orow != nullptr)
            copy_row(reinterpret_cast<const char*>(W) + (size_t)bucket * D * (w_dtype == OOV_F32 ? 4 : 2), w_dtype,
                     orow, out_dtype, D, lane, 32);
    }
    if (tie_count != nullptr && lane == 0 && my_ties) atomicAdd(tie_count, (unsigned long long)my_ties);
}

// lane = id variant (bits_req <= NP planes): the warp-per-id kernel above keeps 10 of 32 lanes busy and walks a
// dependent chain per id (623 us per 1 M ids = 10-16 % of the HBM rate of its algorithmic bytes).  Here every lane
// projects ITS id's feature row on all planes (plane values are warp-uniform shared-memory broadcasts, the row comes
// in 16-byte loads), same f-ascending fmaf chain per plane as above -> identical bits; the bucket rows are then copied
// by the whole warp, one list position after the other, so the stores are full lines.
template <int NP>
__global__ void __launch_bounds__(SLSH_THREADS)
slsh_embed_lane(const float* __restrict__ feat, int64_t n_feat_rows, int F,
                const float* __restrict__ planes, int bits_req, int n_buckets,
                const void* __restrict__ W, int w_dtype,
                const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n,
                int64_t n_old, int64_t prime_pad,
                const void* __restrict__ iv_table, int iv_dtype,
                void* __restrict__ out, int out_dtype, int64_t out_stride, int D,
                float tie_eps, int64_t* __restrict__ bucket_out, unsigned long long* __restrict__ tie_count) {
    extern __shared__ float Pt[];            // [F][NP], plane p of feature f at Pt[f * NP + p]; planes >= bits_req are 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < F * NP; i += SLSH_THREADS) {
        const int f = i / NP, p = i - f * NP;
        Pt[i] = p < bits_req ? __ldg(planes + (size_t)p * F + f) : 0.f;
    }
    __syncthreads();
    const int warps_per_block = SLSH_THREADS / 32;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2, wsz = w_dtype == OOV_F32 ? 4 : 2, isz = iv_dtype == OOV_F32 ? 4 : 2;
    const bool vec4 = (F & 3) == 0 && aligned_dev(feat, 16);
    unsigned int my_ties = 0;
    for (int64_t base = ((int64_t)blockIdx.x * warps_per_block + warp) * 32; base < n;
         base += (int64_t)gridDim.x * warps_per_block * 32) {
        const int64_t r = base + lane;
        const bool valid = r < n;
        const int64_t id = valid ? ids[r * ids_stride] : 0;
        const bool oov = valid && id >= n_old;
        const int64_t fr = feature_row(id, prime_pad);
        const bool inrange = oov && fr >= 0 && fr < n_feat_rows;
        const float* x = feat + (inrange ? fr : 0) * F;
        float acc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] = 0.f;
        if (vec4) {
            for (int f = 0; f < F; f += 4) {
                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inrange) xv = __ldg(reinterpret_cast<const float4*>(x + f));
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4* pr = reinterpret_cast<const float4*>(Pt + (f + i) * NP);
#pragma unroll
                    for (int p4 = 0; p4 < NP / 4; ++p4) {
                        const float4 pv = pr[p4];
                        acc[4 * p4 + 0] = fmaf(xs[i], pv.x, acc[4 * p4 + 0]);
                        acc[4 * p4 + 1] = fmaf(xs[i], pv.y, acc[4 * p4 + 1]);
                        acc[4 * p4 + 2] = fmaf(xs[i], pv.z, acc[4 * p4 + 2]);
                        acc[4 * p4 + 3] = fmaf(xs[i], pv.w, acc[4 * p4 + 3]);
                    }
                }
            }
        } else {
            for (int f = 0; f < F; ++f) {
                const float xf = inrange ? __ldg(x + f) : 0.f;
#pragma unroll
                for (int p = 0; p < NP; ++p) acc[p] = fmaf(xf, Pt[f * NP + p], acc[p]);
            }
        }
        unsigned word = 0u;
        int ties = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (p < bits_req) {
                word |= (!(acc[p] < 0.f)) ? (1u << p) : 0u;
                ties += (fabsf(acc[p]) < tie_eps) ? 1 : 0;
            }
        }
        if (oov) my_ties += (unsigned)ties;
        // (2 ** H).sum(1) % n_buckets with H in {0,1}  ==  (bits_req + popcount) % n_buckets
        const int bucket = (bits_req + __popc(word)) % n_buckets;
        if (bucket_out && valid) bucket_out[r] = oov ? bucket : -1;
        if (out == nullptr) continue;
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        // fast copy: bucket table and output in the same dtype, rows of 64 / 128 / 256 / 512 bytes, nothing in-vocab in
        // this group of 32: LPR lanes move one row in 16-byte pieces, 32 / LPR rows per step, four steps of loads in
        // flight before their stores
        const int row_bytes = D * (int)osz;
        const bool fast_rows = W != nullptr && w_dtype == out_dtype && out_stride == D &&
                               (row_bytes == 64 || row_bytes == 128 || row_bytes == 256 || row_bytes == 512) &&
                               aligned_dev(W, 16) && aligned_dev(out, 16) && __all_sync(0xffffffffu, oov || !valid);
        if (fast_rows) {
            const int lpr = row_bytes >> 4, rpi = 32 / lpr;           // lanes per row, rows per step
            const int piece = lane % lpr, slot = lane / lpr;
            char* obase = reinterpret_cast<char*>(out) + (size_t)base * row_bytes;
            for (int j0 = 0; j0 < cnt; j0 += 4 * rpi) {
                uint4 v[4];
                int jj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    jj[u] = j0 + u * rpi + slot;
                    const int bj = __shfl_sync(0xffffffffu, bucket, jj[u] & 31);
                    if (jj[u] < cnt) v[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(W) + (size_t)bj * row_bytes) + piece);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (jj[u] < cnt) reinterpret_cast<uint4*>(obase + (size_t)jj[u] * row_bytes)[piece] = v[u];
            }
            continue;
        }
        for (int j = 0; j < cnt; ++j) {
            const int64_t idj = __shfl_sync(0xffffffffu, id, j);
            const int oovj = __shfl_sync(0xffffffffu, (int)oov, j), bj = __shfl_sync(0xffffffffu, bucket, j);
            char* orow = reinterpret_cast<char*>(out) + (size_t)(base + j) * out_stride * osz;
            if (oovj) {
                if (W != nullptr)
                    copy_row(reinterpret_cast<const char*>(W) + (size_t)bj * D * wsz, w_dtype, orow, out_dtype, D, lane, 32);
            } else if (iv_table != nullptr && idj >= 0) {
                copy_row(reinterpret_cast<const char*>(iv_table) + (size_t)idj * D * isz, iv_dtype, orow, out_dtype, D, lane, 32);
            }
        }
    }
    my_ties = __reduce_add_sync(0xffffffffu, my_ties);
    if (tie_count != nullptr && lane == 0 && my_ties) atomicAdd(tie_count, (unsigned long long)my_ties);
}

// bucket ids from packed words (F > SLSH_MAXF fallback): one thread per row
__global__ void slsh_finish(const uint32_t* __restrict__ bits, int bits_req, int n_buckets,
                            const void* __restrict__ W, int w_dtype,
                            const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n, int64_t n_old,
                            const void* __restrict__ iv_table, int iv_dtype,
                            void* __restrict__ out, int out_dtype, int64_t out_stride, int D,
                            int64_t* __restrict__ bucket_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2;
    for (int64_t r = warp_global; r < n; r += nwarps) {
        const int64_t id = ids[r * ids_stride];
        char* orow = out ? reinterpret_cast<char*>(out) + (size_t)r * out_stride * osz : nullptr;
        if (id < n_old) {
            if (bucket_out && lane == 0) bucket_out[r] = -1;
            if (iv_table != nullptr && orow != nullptr && id >= 0)
                copy_row(reinterpret_cast<const char*>(iv_table) + (size_t)id * D * (iv_dtype == OOV_F32 ? 4 : 2), iv_dtype,
                         orow, out_dtype, D, lane, 32);
            continue;
        }
        const uint32_t mask = bits_req >= 32 ? 0xffffffffu : ((1u << bits_req) - 1u);
        const int bucket = (bits_req + __popc(bits[r] & mask)) % n_buckets;
        if (bucket_out && lane == 0) bucket_out[r] = bucket;
        if (W != nullptr && orow != nullptr)
            copy_row(reinterpret_cast<const char*>(W) + (size_t)bucket * D * (w_dtype == OOV_F32 ? 4 : 2), w_dtype,
                     orow, out_dtype, D, lane, 32);
    }
}

static int check_rows(const oov_rows* r, const char* who) {
    OOV_REQUIRE(r != nullptr, OOV_ERR_ARG, "%s: rows is NULL", who);
    OOV_REQUIRE(r->n >= 0 && r->D > 0, OOV_ERR_ARG, "%s: bad n=%lld D=%d", who, (long long)r->n, r->D);
    OOV_REQUIRE(r->n == 0 || r->ids != nullptr, OOV_ERR_ARG, "%s: ids is NULL", who);
    OOV_REQUIRE(r->ids_stride >= 1, OOV_ERR_ARG, "%s: ids_stride must be >= 1", who);
    OOV_REQUIRE(dtype_ok(r->out_dtype) && (r->iv_table == nullptr || dtype_ok(r->iv_dtype)), OOV_ERR_ARG, "%s: bad dtype", who);
    OOV_REQUIRE(r->n == 0 || r->out != nullptr, OOV_ERR_ARG, "%s: out is NULL", who);
    OOV_REQUIRE(r->out_stride >= r->D, OOV_ERR_ARG, "%s: out_stride < D", who);
    OOV_REQUIRE(r->n_old >= 0 && r->prime_pad >= 0, OOV_ERR_ARG, "%s: negative n_old / prime_pad", who);
    return OOV_OK;
}

int launch_lsh_bits_simt(const float* feat, int64_t n_feat_rows, int F, const float* planes, int B,
                         const int64_t* ids, int64_t ids_stride, int64_t n, int64_t n_old, int64_t prime_pad,
                         float tie_eps, uint32_t* bits, unsigned long long* tie_count, cudaStream_t st) {
    if (n == 0) return OOV_OK;
    const int words = (B + 31) / 32;
    lsh_bits_simt<<<(unsigned)cdiv(n, BITS_TM), BITS_THREADS, 0, st>>>(feat, n_feat_rows, F, planes, B, ids, ids_stride, n,
                                                                      n_old, prime_pad, tie_eps, bits, words, tie_count);
    OOV_LAUNCH_CHECK("lsh_bits_simt");
    return OOV_OK;
}

int check_rows_public(const oov_rows* r, const char* who) { return check_rows(r, who); }

namespace tc {
bool lsh_tc_supported(int F, int B, int D);
size_t lsh_tc_workspace(int64_t n, int B);
int launch_cast_f32_bf16(const float* src, void* dst, int64_t n_elems, cudaStream_t st);
int lsh_tc_run(const float* feat, int64_t n_feat_rows, int F, const float* planes, int B, const void* W, int w_dtype,
               const oov_rows* rows, float tie_eps, uint32_t* bits_out, unsigned long long* tie_count, void* workspace,
               size_t workspace_bytes, cudaStream_t st, const float* cast_src = nullptr, void* cast_dst = nullptr,
               int64_t cast_elems = 0);
}  // namespace tc

}  // namespace oov

using namespace oov;

extern "C" {

int oov_lsh_bits(const float* feat, int64_t n_feat_rows, int32_t F, const float* planes, int32_t B,
                 const int64_t* ids, int64_t ids_stride, int64_t n, int64_t prime_pad,
                 float tie_eps, uint32_t* bits, unsigned long long* tie_count, int32_t path, void* stream) {
    OOV_REQUIRE(feat && planes && (n == 0 || (ids && bits)), OOV_ERR_ARG, "oov_lsh_bits: NULL pointer");
    OOV_REQUIRE(F > 0 && B > 0 && n >= 0 && n_feat_rows > 0 && ids_stride >= 1, OOV_ERR_ARG,
                "oov_lsh_bits: bad shape F=%d B=%d n=%lld", F, B, (long long)n);
    OOV_REQUIRE(path == OOV_PATH_AUTO || path == OOV_PATH_SIMT_FP32, OOV_ERR_ARG,
                "oov_lsh_bits: only the fp32 path yields exact-sign bits (path=%d)", path);
    return launch_lsh_bits_simt(feat, n_feat_rows, F, planes, B, ids, ids_stride, n, 0, prime_pad, tie_eps, bits,
                                tie_count, (cudaStream_t)stream);
}

size_t oov_lsh_embed_workspace(int64_t n, int32_t B, int32_t D, int32_t path) {
    const int64_t chunk = n < (1 << 20) ? n : (1 << 20);
    size_t a = align_up((size_t)chunk * ((B + 31) / 32) * 4, 256);
    if (path != OOV_PATH_SIMT_FP32 && B > 0) {                 // F is not known here: reserve for the tensor-core path too
        const size_t b = tc::lsh_tc_workspace(n, B);
        if (b > a) a = b;
    }
    (void)D;
    return a;
}

int oov_lsh_embed(const float* feat, int64_t n_feat_rows, int32_t F, const float* planes, int32_t B,
                  const void* W, int32_t w_dtype, const oov_rows* rows, float tie_eps,
                  uint32_t* bits_out, unsigned long long* tie_count,
                  void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    return oov_lsh_embed_cast(feat, n_feat_rows, F, planes, B, W, w_dtype, rows, tie_eps, bits_out, tie_count, workspace,
                              workspace_bytes, path, nullptr, nullptr, 0, stream);
}

int oov_lsh_embed_cast(const float* feat, int64_t n_feat_rows, int32_t F, const float* planes, int32_t B,
                       const void* W, int32_t w_dtype, const oov_rows* rows, float tie_eps,
                       uint32_t* bits_out, unsigned long long* tie_count,
                       void* workspace, size_t workspace_bytes, int32_t path,
                       const float* cast_src, void* cast_dst, int64_t cast_elems, void* stream) {
    int rc = check_rows(rows, "oov_lsh_embed");
    if (rc) return rc;
    OOV_REQUIRE(cast_elems >= 0 && cast_elems % 8 == 0, OOV_ERR_ARG, "oov_lsh_embed_cast: cast_elems=%lld must be a multiple of 8", (long long)cast_elems);
    OOV_REQUIRE(cast_elems == 0 || (cast_src && cast_dst && aligned(cast_src, 16) && aligned(cast_dst, 16)), OOV_ERR_ALIGN,
                "oov_lsh_embed_cast: cast_src / cast_dst must be non-NULL and 16-byte aligned");
    OOV_REQUIRE(feat && planes && W && dtype_ok(w_dtype), OOV_ERR_ARG, "oov_lsh_embed: NULL pointer / bad dtype");
    OOV_REQUIRE(F > 0 && B > 0 && n_feat_rows > 0, OOV_ERR_ARG, "oov_lsh_embed: bad shape F=%d B=%d", F, B);
    OOV_REQUIRE(path >= OOV_PATH_AUTO && path <= OOV_PATH_TCGEN05, OOV_ERR_ARG, "oov_lsh_embed: unsupported path %d", path);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows->n == 0) return tc::launch_cast_f32_bf16(cast_src, cast_dst, cast_elems, st);
    // tensor cores (split-bf16 exact-sign GEMM fused with the bucket-mean GEMM) when the tile shapes allow
    const bool tc_ok = tc::lsh_tc_supported(F, B, rows->D);
    OOV_REQUIRE(path != OOV_PATH_TCGEN05 || tc_ok, OOV_ERR_ARG, "oov_lsh_embed: tcgen05 path needs F <= 64 and D <= 64 (F=%d D=%d)", F, rows->D);
    if (tc_ok && path != OOV_PATH_SIMT_FP32)
        return tc::lsh_tc_run(feat, n_feat_rows, F, planes, B, W, w_dtype, rows, tie_eps, bits_out, tie_count, workspace,
                              workspace_bytes, st, cast_src, cast_dst, cast_elems);
    rc = tc::launch_cast_f32_bf16(cast_src, cast_dst, cast_elems, st);
    if (rc) return rc;
    const int words = (B + 31) / 32;
    const int64_t chunk = bits_out ? rows->n : (rows->n < (1 << 20) ? rows->n : (1 << 20));
    if (!bits_out) {
        OOV_REQUIRE(workspace && workspace_bytes >= (size_t)chunk * words * 4, OOV_ERR_WORKSPACE,
                    "oov_lsh_embed: workspace %zu < %zu", workspace_bytes, (size_t)chunk * words * 4);
    }
    const size_t osz = dtype_size(rows->out_dtype);
    for (int64_t r0 = 0; r0 < rows->n; r0 += chunk) {
        const int64_t cn = rows->n - r0 < chunk ? rows->n - r0 : chunk;
        uint32_t* bits = bits_out ? bits_out + r0 * words : reinterpret_cast<uint32_t*>(workspace);
        const int64_t* ids = rows->ids + r0 * rows->ids_stride;
        rc = launch_lsh_bits_simt(feat, n_feat_rows, F, planes, B, ids, rows->ids_stride, cn, rows->n_old,
                                  rows->prime_pad, tie_eps, bits, tie_count, st);
        if (rc) return rc;
        lsh_mean_simt<<<(unsigned)cdiv(cn, MEAN_TM), MEAN_THREADS, 0, st>>>(
            bits, words, B, W, w_dtype, rows->D, ids, rows->ids_stride, cn, rows->n_old, rows->iv_table, rows->iv_dtype,
            reinterpret_cast<char*>(rows->out) + (size_t)r0 * rows->out_stride * osz, rows->out_dtype, rows->out_stride);
        OOV_LAUNCH_CHECK("lsh_mean_simt");
    }
    return OOV_OK;
}

int oov_slsh_embed(const float* feat, int64_t n_feat_rows, int32_t F, const float* planes, int32_t bits_req,
                   int32_t n_buckets, const void* W, int32_t w_dtype, const oov_rows* rows, float tie_eps,
                   int64_t* bucket_out, unsigned long long* tie_count, void* stream) {
    OOV_REQUIRE(rows != nullptr, OOV_ERR_ARG, "oov_slsh_embed: rows is NULL");
    if (W != nullptr || rows->out != nullptr) {
        int rc = check_rows(rows, "oov_slsh_embed");
        if (rc) return rc;
    }
    OOV_REQUIRE(feat && planes, OOV_ERR_ARG, "oov_slsh_embed: NULL pointer");
    OOV_REQUIRE(W == nullptr || dtype_ok(w_dtype), OOV_ERR_ARG, "oov_slsh_embed: bad w_dtype");
    OOV_REQUIRE(F > 0 && n_feat_rows > 0 && n_buckets > 0, OOV_ERR_ARG, "oov_slsh_embed: bad shape");
    OOV_REQUIRE(bits_req >= 0 && bits_req <= 32, OOV_ERR_ARG, "oov_slsh_embed: bits_req=%d not in [0,32]", bits_req);
    if (rows->n == 0) return OOV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    OOV_REQUIRE(F <= SLSH_MAXF, OOV_ERR_ARG, "oov_slsh_embed: F=%d > %d (use oov_lsh_bits + bucket gather)", F, SLSH_MAXF);
    if (bits_req <= 16 && F <= 512) {
        constexpr int NP = 16;
        const size_t smem16 = (size_t)F * NP * sizeof(float);
        // (function attributes are per device: set on every call — cheap and idempotent; occupancy cached per device)
        cudaFuncSetAttribute(slsh_embed_lane<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * NP * 4);
        static int per_sm_dev[64] = {0};
        int& per_sm = per_sm_dev[cur_device()];
        if (per_sm == 0 &&
            (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slsh_embed_lane<NP>, SLSH_THREADS, smem16) != cudaSuccess || per_sm < 1))
            per_sm = 2;
        int64_t blocks16 = cdiv(rows->n, (int64_t)(SLSH_THREADS / 32) * 32);
        if (blocks16 > (int64_t)num_sms() * per_sm) blocks16 = (int64_t)num_sms() * per_sm;
        slsh_embed_lane<NP><<<(unsigned)blocks16, SLSH_THREADS, smem16, st>>>(
            feat, n_feat_rows, F, planes, bits_req, n_buckets, W, w_dtype, rows->ids, rows->ids_stride, rows->n, rows->n_old,
            rows->prime_pad, rows->iv_table, rows->iv_dtype, rows->out, rows->out_dtype, rows->out_stride, rows->D, tie_eps,
            bucket_out, tie_count);
        OOV_LAUNCH_CHECK("slsh_embed_lane");
        return OOV_OK;
    }
    const size_t smem = (size_t)F * 32 * sizeof(float);
    cudaFuncSetAttribute(slsh_embed_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, SLSH_MAXF * 32 * 4);
    const int warps_per_block = SLSH_THREADS / 32;
    int64_t blocks = cdiv(rows->n, warps_per_block);
    const int64_t max_blocks = (int64_t)num_sms() * (smem > 64 * 1024 ? 1 : (smem > 24 * 1024 ? 2 : 8));
    if (blocks > max_blocks) blocks = max_blocks;
    slsh_embed_simt<<<(unsigned)blocks, SLSH_THREADS, smem, st>>>(
        feat, n_feat_rows, F, planes, bits_req, n_buckets, W, w_dtype, rows->ids, rows->ids_stride, rows->n, rows->n_old,
        rows->prime_pad, rows->iv_table, rows->iv_dtype, rows->out, rows->out_dtype, rows->out_stride, rows->D, tie_eps,
        bucket_out, tie_count);
    OOV_LAUNCH_CHECK("slsh_embed_simt");
    return OOV_OK;
}

}  // extern "C"
