/*
 * oov_b200.h — C-ABI of the B200-native OOV inductive-embedding + full-sort top-k path.
 *
 * One shared library (liboov_b200.so, sm_100a only, no CPU fallback).  Every entry
 * point is `extern "C"`, takes plain device pointers + sizes (no torch types), is
 * asynchronous on the given CUDA stream (`stream` = cudaStream_t cast to void*),
 * performs no host synchronisation and no persistent allocation, and returns an
 * int status (0 = ok, negative = error; `oov_last_error()` gives the message).
 * The reference (snap-research/improving-inductive-oov-recsys) is pure Python: the
 * "FFI" a maintainer adds is a ctypes binding; each entry point
 * below names the reference code it replaces (paths relative to RecBole/recbole/).
 *
 * Conventions
 *   - dtype codes: OOV_F32 (0) or OOV_BF16 (1) for tables / outputs.
 *   - ids are int64 (oov_prime_pad = 112062759511 > 2^32, properties/overall.yaml:71).
 *   - `oov_rows` describes the "assemble" contract shared by all embed calls: for row i,
 *       id = ids[i * ids_stride]
 *       id <  n_old : out[i] = iv_table[id]      (in-vocab gather; skipped if iv_table == NULL)
 *       id >= n_old : out[i] = embedder(id)      (OOV)
 *     which is model/general_recommender/bpr.py:48-125 / directau.py:107-172 without the
 *     boolean-mask indexing (and its device->host syncs).
 */
#ifndef OOV_B200_H_
#define OOV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OOV_OK 0
#define OOV_ERR_ARG (-1)          /* bad shape / null pointer / unsupported size */
#define OOV_ERR_ALIGN (-2)        /* pointer or stride not aligned as required   */
#define OOV_ERR_ARCH (-3)         /* device is not sm_100                        */
#define OOV_ERR_CUDA (-4)         /* a CUDA runtime call failed                  */
#define OOV_ERR_WORKSPACE (-5)    /* workspace too small                         */

#define OOV_F32 0
#define OOV_BF16 1

/* precision / path selector for the GEMM-shaped ops */
#define OOV_PATH_AUTO 0           /* tcgen05 where the shape allows, else SIMT   */
#define OOV_PATH_SIMT_FP32 1      /* CUDA-core fp32 FMA (exact-sign projections) */
#define OOV_PATH_TCGEN05 2        /* tensor cores (tcgen05.mma + TMEM)           */

const char* oov_version(void);
const char* oov_last_error(void);
/* OOV_OK iff `device` is compute capability 10.x; the Python host refuses to run otherwise. */
int oov_check_device(int device);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
uint64_t oov_launch_count(void);

typedef struct oov_rows {
    const int64_t* ids;     /* device; n ids spaced ids_stride elements apart            */
    int64_t ids_stride;     /* 1 = contiguous; `fields` for a column of a [B,fields] batch */
    int64_t n;
    int64_t n_old;          /* ids < n_old are in-vocab                                  */
    int64_t prime_pad;      /* 0 = eval; > 0 = training mode: feature row = id - prime_pad
                               when id >= prime_pad (inductive/lsh_embedder.py:153-155)   */
    const void* iv_table;   /* [>= n_old, D]; NULL = leave in-vocab rows of out untouched */
    int32_t iv_dtype;
    int32_t out_dtype;
    void* out;              /* row i at out + i * out_stride elements                    */
    int64_t out_stride;
    int32_t D;
    int32_t _pad;
} oov_rows;

/* ------------------------------------------------------------------------------------
 * LSH  — replaces inductive/torch_hash.py:55-60 (TorchLSHash.hash_points) and
 *        inductive/lsh_embedder.py:116-179 (_hash_node, embed_user_ids, embed_item_ids).
 * feat   : [n_feat_rows, F] fp32 row-major feature matrix (lsh_embedder.py:80-106)
 * planes : [B, F] fp32 (uniform_planes[0]); B = n_oov_buckets for `lsh`
 * bits   : uint32 [n, ceil(B/32)], bit (b & 31) of word (b >> 5) = !(x.p_b < 0)
 *          (so +0, -0 and NaN give 1, as in torch_hash.py:57-59)
 * tie_count : optional device counter, incremented once per projection with |x| < tie_eps
 * ------------------------------------------------------------------------------------ */
int oov_lsh_bits(const float* feat, int64_t n_feat_rows, int32_t F,
                 const float* planes, int32_t B,
                 const int64_t* ids, int64_t ids_stride, int64_t n, int64_t prime_pad,
                 float tie_eps, uint32_t* bits, unsigned long long* tie_count,
                 int32_t path, void* stream);

/* out[i] = (H_i @ W) / sum(H_i) for OOV rows — lsh_embedder.py:156-158,176-178; an
 * all-zero H_i gives NaN (0/0) like the reference.  `bits_out` (optional) receives the
 * multi-hot words of every row (garbage for in-vocab rows). */
int oov_lsh_embed(const float* feat, int64_t n_feat_rows, int32_t F,
                  const float* planes, int32_t B,
                  const void* oov_weight, int32_t w_dtype,     /* [B, D] model.*_oov_buckets.weight */
                  const oov_rows* rows, float tie_eps,
                  uint32_t* bits_out, unsigned long long* tie_count,
                  void* workspace, size_t workspace_bytes,
                  int32_t path, void* stream);
size_t oov_lsh_embed_workspace(int64_t n, int32_t B, int32_t D, int32_t path);
/* oov_lsh_embed plus a side job: cast_dst[0 .. cast_elems) = bf16(cast_src[0 .. cast_elems)) (cast_elems % 8 == 0, both
 * 16-byte aligned) — the in-vocab half of the same bf16 item table when its ids are a contiguous range
 * (bpr.py:111-112 `item_e[in_vocab_items] = item_embedding(item[in_vocab_items])` for item = arange).  On the tcgen05
 * path with enough OOV rows to fill the machine the copy is done by the TMA warp of every CTA while it waits for a free
 * ring stage, in the shadow of the GEMM pipeline; otherwise it is a separate launch on the same stream. */
int oov_lsh_embed_cast(const float* feat, int64_t n_feat_rows, int32_t F,
                       const float* planes, int32_t B,
                       const void* oov_weight, int32_t w_dtype,
                       const oov_rows* rows, float tie_eps,
                       uint32_t* bits_out, unsigned long long* tie_count,
                       void* workspace, size_t workspace_bytes, int32_t path,
                       const float* cast_src, void* cast_dst, int64_t cast_elems, void* stream);

/* ------------------------------------------------------------------------------------
 * SLSH — replaces inductive/single_lsh_embedder.py:77-109.
 * planes [bits_req, F]; bucket = (bits_req + popcount(bits)) % n_buckets
 * (== ((2 ** H).sum(1)).long() % n_buckets, single_lsh_embedder.py:86); out[i] = W[bucket].
 * bucket_out (optional int64 [n]) receives the bucket id of every OOV row (-1 for in-vocab).
 * ------------------------------------------------------------------------------------ */
int oov_slsh_embed(const float* feat, int64_t n_feat_rows, int32_t F,
                   const float* planes, int32_t bits_req, int32_t n_buckets,
                   const void* oov_weight, int32_t w_dtype,    /* [n_buckets, D]; NULL = ids only */
                   const oov_rows* rows, float tie_eps,
                   int64_t* bucket_out, unsigned long long* tie_count, void* stream);

/* ------------------------------------------------------------------------------------
 * DHE — replaces inductive/dh_embedder.py:140-170 (_get_hashes/_hash_ids; csiphash
 *       SipHash-2-4) and :70-89,191-217 (the 4-layer hash nets).
 * keys   : device uint8 [H, 16]
 * hashes : uint32 [n, H] = LE_u64(SipHash-2-4(key_j, LE8(id_i))) % mod   (mod = 2^24)
 * ------------------------------------------------------------------------------------ */
int oov_dhe_hash(const int64_t* ids, int64_t ids_stride, int64_t n,
                 const uint8_t* keys, int32_t H, uint64_t mod,
                 uint32_t* hashes, void* stream);

typedef struct oov_dhe_net {
    const float* w[4];      /* nn.Linear weights [out, in]: [hid,H+F], [hid,hid], [hid,hid], [D,hid] */
    const float* b[4];
    int32_t H, hidden, D;
    int32_t F;              /* plain fp32 feature inputs after the H hash inputs of layer 1: 0 for `dhe`;
                               `fdhe` H + F, `dnn` H = 0 (oov_fdhe_embed only)                        */
} oov_dhe_net;

/* out[i] = Sigmoid(L4(GELU(L3(GELU(L2(GELU(L1(float(hashes_i)))))))))  (erf GELU) */
int oov_dhe_mlp(const uint32_t* hashes, int64_t n, const oov_dhe_net* net,
                void* out, int32_t out_dtype, int64_t out_stride,
                void* workspace, size_t workspace_bytes, int32_t path, void* stream);
/* hash + MLP + assemble (dh_embedder.py:219-245; ids are NOT de-padded there). */
int oov_dhe_embed(const uint8_t* keys, uint64_t mod, const oov_dhe_net* net,
                  const oov_rows* rows, void* workspace, size_t workspace_bytes,
                  int32_t path, void* stream);
size_t oov_dhe_workspace(int64_t n, const oov_dhe_net* net, int32_t path);
/* The reference memoises the hashes of an id (dh_embedder.py:139 `@cache` on _get_hashes): they depend on (id, keys)
 * only, never on the weights.  oov_dhe_hash_planes writes them once in the layout the tensor-core MLP consumes — three
 * exact bf16 byte planes per id, h = 65536 a + 256 b + c: planes[i, j] = a, planes[i, H + j] = b, planes[i, 2H + j] = c,
 * row stride oov_dhe_planes_ld(H) elements (padding columns zero) — and oov_dhe_embed_planes runs the MLP + assemble
 * from the memoised planes (rows->ids is only read for the in-vocab / OOV decision when rows->n_old > 0).
 * mod must be <= 2^24; tcgen05 path only. */
int64_t oov_dhe_planes_ld(int32_t H);
int oov_dhe_hash_planes(const int64_t* ids, int64_t ids_stride, int64_t n,
                        const uint8_t* keys, int32_t H, uint64_t mod,
                        void* planes /* bf16 [n, ld] */, void* stream);
int oov_dhe_embed_planes(const void* planes, const oov_dhe_net* net, const oov_rows* rows,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * fdhe / dnn — replaces inductive/feat_dh_embedder.py:188-196 (_hash_users/_hash_items: hstack(hashes, feature row)
 *              -> 4-layer net) and inductive/dnn_embedder.py:87-91 (feature row -> the same net without hashes).
 * net->w[0] is [hidden, H + F]: columns [0, H) take the raw hash values (as fp32, like `dhe`), columns [H, H + F) the
 * feature row feat[id'] with id' = id - prime_pad when rows->prime_pad > 0 and id >= prime_pad
 * (feat_dh_embedder.py:199-205: the hashes use the ORIGINAL id, only the feature lookup is de-padded); a row id'
 * outside [0, n_feat_rows) reads zeros.  net->H == 0 (keys may be NULL): `dnn`.  The tcgen05 path rounds features and
 * weights to bf16 (hash inputs stay exact, three byte planes), fp32 accumulate; OOV_PATH_SIMT_FP32 is all fp32.
 * workspace: oov_fdhe_workspace(n, net, path) bytes.
 * ------------------------------------------------------------------------------------ */
int oov_fdhe_embed(const uint8_t* keys, uint64_t mod, const oov_dhe_net* net,
                   const float* feat, int64_t n_feat_rows,
                   const oov_rows* rows, void* workspace, size_t workspace_bytes,
                   int32_t path, void* stream);
size_t oov_fdhe_workspace(int64_t n, const oov_dhe_net* net, int32_t path);

/* One bf16 linear layer on the tensor cores (the building block of the tcgen05 DHE path):
 * out[M, N] = act(A[M, K] . W[N, K]^T + bias); A, W bf16 row-major (lda, ldw in elements, multiples of 8),
 * fp32 accumulate in TMEM; act: 0 none, 1 GELU(erf), 2 sigmoid; out fp32 or bf16. */
int oov_tc_linear(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int32_t N, int32_t K,
                  const float* bias, int32_t act, void* out, int32_t out_dtype, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------
 * mean / zero — replaces inductive/mean_embedder.py:42-87, zero_embedder.py:36-60.
 * ------------------------------------------------------------------------------------ */
/* mean[d] = (1/rows) * sum_r table[r, d]  (ALL rows, incl. the pad row 0); fp32 out.
 * workspace: oov_col_mean_workspace(rows, D) bytes. */
int oov_col_mean(const void* table, int32_t dtype, int64_t rows, int32_t D,
                 float* mean_out, void* workspace, size_t workspace_bytes, void* stream);
size_t oov_col_mean_workspace(int64_t rows, int32_t D);
/* OOV rows get the constant vector `vec` (fp32 [D]; NULL = zeros), in-vocab rows are gathered. */
int oov_const_embed(const float* vec, const oov_rows* rows, void* stream);

/* Plain row gather out[i] = table[idx[i]] (nn.Embedding lookups bpr.py:80-84; the
 * inductive_mapper path's *_oov_buckets(mapped_id - n_old)). */
int oov_gather_rows(const void* table, int32_t dtype, int64_t table_rows, int32_t D,
                    const int64_t* idx, int64_t idx_stride, int64_t n, int64_t idx_offset,
                    void* out, int32_t out_dtype, int64_t out_stride, void* stream);

/* ------------------------------------------------------------------------------------
 * Full-sort scoring + masks + top-k — replaces bpr.py:151-156 / directau.py:193-198
 * (score = user_e @ all_item_e.T), inductive/evaluator.py:91-94 (pad column + history
 * -> -inf) and evaluator/collector.py:153-159 (torch.topk(scores, max(topk))).
 *
 * users [Q, D], items [N, D] (this shard's rows), same dtype.  Scores are fp32.
 * Masked to -inf: global item id 0 (pad) when mask_pad != 0; ids outside
 * [seg_lo, seg_hi) (inductive/collector_filter.py:172-175; pass 0, INT64_MAX for none);
 * (user q, item) pairs of the CSR history (hist_rowptr int32 [Q+1], hist_cols int32,
 * global item ids, ascending per row; NULL = none).
 * Output per user: k (score, global id) pairs ordered by (score desc, id asc); NaN
 * ranks above every number like torch.topk.  global id = local row + item_id_offset.
 * If N < k the tail is filled with (-inf, -1).
 * ------------------------------------------------------------------------------------ */
int oov_fullsort_topk(const void* users, const void* items, int32_t dtype,
                      int64_t Q, int64_t N, int32_t D, int32_t k,
                      int64_t item_id_offset, int32_t mask_pad,
                      int64_t seg_lo, int64_t seg_hi,
                      const int32_t* hist_rowptr, const int32_t* hist_cols,
                      float* out_scores, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, int32_t path, void* stream);
size_t oov_fullsort_topk_workspace(int64_t Q, int64_t N, int32_t D, int32_t k, int32_t path);

/* Dense scores [Q, N] fp32 (the reference's materialised matrix; kept for parity tests
 * and for callers that need `rec.score`): masks applied as above. */
int oov_fullsort_scores(const void* users, const void* items, int32_t dtype,
                        int64_t Q, int64_t N, int32_t D,
                        int64_t item_id_offset, int32_t mask_pad, int64_t seg_lo, int64_t seg_hi,
                        const int32_t* hist_rowptr, const int32_t* hist_cols,
                        float* scores, int64_t scores_stride, void* stream);

/* torch.topk(scores, k) of an already materialised [Q, N] fp32 matrix (collector.py:153-159 for
 * callers that hold dense scores): (score desc, index asc), NaN first. */
int oov_dense_topk(const float* scores, int64_t scores_stride, int64_t Q, int64_t N, int32_t k,
                   float* out_scores, int64_t* out_idx, void* stream);

/* Merge G per-shard candidate lists [G, Q, k] into the global top-k (the
 * lists arrive by NCCL all-gather).  Order: (score desc, id asc), NaN first. */
int oov_topk_merge(const float* cand_scores, const int64_t* cand_idx,
                   int32_t G, int64_t Q, int32_t k,
                   float* out_scores, int64_t* out_idx, void* stream);

/* Row-sharded exchange with 8-byte candidates: oov_fullsort_topk_keys is oov_fullsort_topk for a shard
 * table laid out [global ids lo0 .. lo0 + n0 | global ids lo1 ..] (one slice of the in-vocab ids, one of the OOV ids)
 * that writes every list entry as ONE packed key in GLOBAL item ids, keys_out [Q, k] uint64 =
 * (order-preserving score bits << 32) | ~uint32(global id), 0 = empty slot, already ordered (score desc, id asc): what
 * the ranks all-gather (8·Q·k bytes each).  seg_lo / seg_hi and the history CSR are in LOCAL rows of the shard table.
 * oov_topk_merge_keys merges G such lists [G, Q, k] into the global (scores, ids).  Global ids must fit 32 bits. */
int oov_fullsort_topk_keys(const void* users, const void* items, int32_t dtype,
                           int64_t Q, int64_t N, int32_t D, int32_t k, int32_t mask_pad,
                           int64_t seg_lo, int64_t seg_hi,
                           const int32_t* hist_rowptr, const int32_t* hist_cols,
                           int64_t n0, int64_t lo0, int64_t lo1, uint64_t* keys_out,
                           void* workspace, size_t workspace_bytes, int32_t path, void* stream);
int oov_topk_merge_keys(const uint64_t* keys, int32_t G, int64_t Q, int32_t k,
                        float* out_scores, int64_t* out_idx, void* stream);

/* 'rec.topk' rows of evaluator/collector.py:160-166: hits[q, j] = 1 iff topk_idx[q, j] is a
 * positive of user q (CSR pos_rowptr/pos_cols, ascending), last column = number of positives. */
int oov_topk_hits(const int64_t* topk_idx, int64_t Q, int32_t k,
                  const int32_t* pos_rowptr, const int32_t* pos_cols,
                  int32_t* out_hits /* [Q, k+1] */, void* stream);

/* The seven collectors of inductive/evaluator.py:29-56 (overall, old_users, new_users, old_old, old_new, new_old,
 * new_new — in this order) from the three k-lists of ONE scoring pass over items split at n_old_items (all / old items
 * only / new items only; the "all" list is the merge of the other two).  Replaces collector_filter.py:128-256 +
 * filtered_collector.py:18-80 (seven masked copies of the score matrix, seven top-k) without boolean indexing or host
 * synchronisation: out [7, Q, k + 2] int32 = hit flags of the collector's list | pos_len | keep, where keep = 1 marks
 * the rows the reference's collector would hold (users passing the user filter that own a positive after the item
 * filter; every user for "overall").  user_ids [Q]; positives as in oov_topk_hits.  reference_compat = 1 reproduces two
 * quirks of collector_filter.py (:172-175 blanks the item segment chosen by return_old_USERS; :249-250 shifts new-item
 * positives by -n_old_items while score columns stay global); 0 filters by return_old_items and compares global ids. */
int oov_topk_hits_collectors(const int64_t* idx_all, const int64_t* idx_old, const int64_t* idx_new,
                             int64_t Q, int32_t k, const int64_t* user_ids,
                             int64_t n_old_users, int64_t n_old_items,
                             const int32_t* pos_rowptr, const int32_t* pos_cols,
                             int32_t reference_compat, int32_t* out, void* stream);

/* (row, item) index pairs — the `history_index` / `positive_u, positive_i` tensors FullSortEvalDataLoader.collate_fn
 * yields (data/dataloader/general_dataloader.py:270-292) — to the CSR the kernels above read: rowptr [Q + 1] int32,
 * cols [n_pairs] int32 ascending within a row (duplicates kept).  Pairs whose row is outside [0, Q) are padding and
 * are dropped (cols_out past rowptr[Q] is left untouched).  col_ranges = {lo0, hi0, lo1, hi1} (host array, optional)
 * maps item ids to the LOCAL rows of a two-range row shard — [lo0, hi0) -> 0.., [lo1, hi1) -> (hi0 - lo0).. — and drops
 * every other item: the history of a shard that holds one slice of the in-vocab ids and one of the OOV ids.  One launch, asynchronous, no workspace:
 * n_pairs < 2^31, 1 <= Q <= 2^24 (ceil(Q / 512) CTAs beyond 8192 rows, each streams the pair list). */
int oov_pairs_to_csr(const int64_t* rows, const int64_t* cols, int64_t n_pairs, int64_t Q,
                     const int64_t* col_ranges /* HOST, 4 values or NULL */,
                     int32_t* rowptr_out, int32_t* cols_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Context models — replaces model/abstract_recommender.py:794-842 (embed_token_fields)
 * with model/layers.py:150-153 (FMEmbedding), and model/layers.py:1634-1693.
 * tokens [Bn, fields] int64; offsets [fields] int64; table [V, D].
 * out[b, f] = table[tokens[b,f] + offsets[f]], except column uid_idx (resp. iid_idx) of
 * rows whose id >= n_users (resp. n_items): those get user_const / item_const (fp32 [D])
 * when given, else are left for an oov_*_embed call with ids_stride = fields,
 * out_stride = fields * D.
 * ------------------------------------------------------------------------------------ */
int oov_token_gather(const int64_t* tokens, int64_t Bn, int32_t fields, const int64_t* offsets,
                     const void* table, int32_t dtype, int64_t table_rows, int32_t D,
                     int64_t n_users, int64_t n_items, int32_t uid_idx, int32_t iid_idx,
                     const float* user_const, const float* item_const,
                     void* out, int32_t out_dtype, void* stream);
/* First-order term: out[b] = sum_f table1[tokens[b,f] + offsets[f]] with the OOV user/item
 * entries replaced by oov_user_val[b] / oov_item_val[b] (fp32 [Bn], produced by a D = 1 embed). */
int oov_first_order_sum(const int64_t* tokens, int64_t Bn, int32_t fields, const int64_t* offsets,
                        const float* table1, int64_t table_rows,
                        int64_t n_users, int64_t n_items, int32_t uid_idx, int32_t iid_idx,
                        const float* oov_user_val, const float* oov_item_val,
                        float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Context dense towers — replaces the eval-mode maths of
 * model/context_aware_recommender/dcnv2.py:120-144 (cross network) and model/layers.py:33-92 (MLPLayers with eval-mode
 * BatchNorm folded into the Linear) on top of oov_tc_linear (act: 0 none, 1 GELU, 2 sigmoid, 3 ReLU).
 * oov_cross_update: out = x0 * t + xl elementwise over bf16 tensors of n_elems elements (multiple of 8, 16-byte
 * aligned) — the tail x_{l+1} = x_0 * (W_l x_l + b_l) + x_l of a cross layer, t being the tensor-core linear's output.
 * ------------------------------------------------------------------------------------ */
int oov_cross_update(const void* x0, const void* t, const void* xl, int64_t n_elems, void* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Training-mode OOV path — the backward of the assemble that trainer/trainer.py:1748-1837
 * (_train_oov -> calculate_loss -> bpr.py:48-125 under autograd) differentiates.  The forward is the eval-path
 * `*_embed` call with rows->prime_pad set.  g: fp32 [n, D] gradient of the assembled rows (row stride ldg).
 * oov_scatter_add_rows: dtable[idx[i] + idx_offset] += g[i] for rows that land in [0, rows) — in-vocab rows
 *   (idx = ids, rows = n_old: nn.Embedding backward), mapper rows (idx_offset = -n_old: bpr.py:71) and slsh rows
 *   (idx = the bucket ids oov_slsh_embed returned, -1 for in-vocab: single_lsh_embedder.py:108).
 * oov_lsh_embed_backward: dW[b] += H_ib * g_i / |H_i| for ids >= n_old (lsh_embedder.py:156-158), bits = the multi-hot
 *   words oov_lsh_embed wrote to `bits_out`; an all-zero hash poisons dW with NaN like autograd's 0 * inf.
 * Both ADD into their fp32 output (zero it first); accumulation order is not deterministic (atomics).
 * ------------------------------------------------------------------------------------ */
int oov_scatter_add_rows(const float* g, int64_t ldg, const int64_t* idx, int64_t idx_stride, int64_t n,
                         int64_t idx_offset, int64_t rows, int32_t D, float* dtable, void* stream);
int oov_lsh_embed_backward(const uint32_t* bits, int32_t B, const float* g, int64_t ldg,
                           const int64_t* ids, int64_t ids_stride, int64_t n, int64_t n_old, int32_t D,
                           float* dW, void* workspace, size_t workspace_bytes, void* stream);
size_t oov_lsh_embed_backward_workspace(int64_t n);

/* Hash-net training (dhe / fdhe / dnn; dh_embedder.py:70-89 under autograd), all fp32:
 * oov_fdhe_input: x[i] = [float(hash_0..H-1 of ids[i]) | feat[id'_i]] — the first layer's input, [n, H + F] contiguous
 *   (same id rules as oov_fdhe_embed; H = 0 or F = 0 allowed); workspace oov_fdhe_input_workspace(n, H) bytes.
 * oov_linear_f32: out[M, N] = act(A[M, K] . W[N, K]^T + bias), contiguous operands, act 0 none / 1 GELU(erf) / 2 sigmoid
 *   — layer forward (act 0: the pre-activations the backward needs) and, with transposed operands, the backward
 *   products dA = dZ W and dW = dZ^T A.
 * oov_act: dy == NULL: out = act(z); else out = dy * act'(z), rows whose id < n_old zeroed when ids != NULL (in-vocab
 *   rows of an assembled batch do not reach the net). */
int oov_fdhe_input(const uint8_t* keys, uint64_t mod, int32_t H, const float* feat, int64_t n_feat_rows, int32_t F,
                   const int64_t* ids, int64_t ids_stride, int64_t n, int64_t prime_pad, float* x,
                   void* workspace, size_t workspace_bytes, void* stream);
size_t oov_fdhe_input_workspace(int64_t n, int32_t H);
int oov_linear_f32(const float* A, const float* W, int64_t M, int32_t N, int32_t K, const float* bias, int32_t act,
                   float* out, void* stream);
int oov_act(const float* z, const float* dy, int32_t act, int64_t rows, int32_t N,
            const int64_t* ids, int64_t ids_stride, int64_t n_old, float* out, void* stream);

/* xDeepFM compressed interaction network (xdeepfm.py:134-190) around oov_tc_linear.  All operands
 * bf16; rows of z and of every CIN layer output are (b, d) pairs (b-major), channels run along the row.
 * oov_cin_outer: z[(b*D + d) * ldz + h*M + m] = xi[b, d, h] * x0[b, d, m] (fp32 product, rounded once), channels
 * [H*M, ldz) zero — the einsum "bhd,bmd->bhmd" + view of xdeepfm.py:160-165 as the A operand of the layer's
 * kernel-size-1 Conv1d.  Element (b, d, c) of xi sits at xi[b*xi_sb + d*xi_sd + c*xi_sc] (elements), same for x0:
 * the gathered embeddings [B, M, D] are read in place (sb = M*D, sd = 1, sc = D), a previous layer's output
 * [B*D, ld] with (sb = D*ld, sd = ld, sc = 1).
 * oov_cin_pool_dot: acc[b] = (accumulate ? acc[b] : 0) + bias + sum_{d, c < ncols} y[(b*D + d)*ldy + col0 + c] * w[c]
 * — the sum pooling over the embedding axis (xdeepfm.py:188-189) of one layer's direct-connect channels folded with
 * that layer's slice of cin_linear (xdeepfm.py:198), fp32. */
int oov_cin_outer(const void* xi, int64_t xi_sb, int64_t xi_sd, int64_t xi_sc, int32_t H,
                  const void* x0, int64_t x0_sb, int64_t x0_sd, int64_t x0_sc, int32_t M,
                  int64_t B, int32_t D, void* z, int64_t ldz, void* stream);
int oov_cin_pool_dot(const void* y, int64_t ldy, int32_t col0, int32_t ncols, int64_t B, int32_t D,
                     const float* w, float bias, int32_t accumulate, float* acc, void* stream);
/* One CIN layer in ONE tcgen05 kernel: the outer-product operand z is generated tile by tile in shared memory (never
 * written to HBM), Y = ReLU(z W^T + bias), then
 *   hid_out[(b*D + d) * ld_h + c] = bf16(Y[., c]) for c < n_hidden   (the next layer's xi: sb = D*ld_h, sd = ld_h, sc = 1)
 *   out_acc[b] += sum_{d, c < pool_n} bf16(Y[(b, d), pool_lo + c]) * pool_w[c]     (fp32 atomics; initialise out_acc with
 *                                                                                   cin_linear's bias)
 * xi [B*D, ld_xi] / x0 [B*D, ld_x0]: rows are (b, d) pairs, channels contiguous (the embeddings transposed to [B, D, M]
 * once per forward; a previous layer's hid_out as it is), even ld, 4-byte aligned.
 * W bf16 [O, ldw]: column h*Mp + m holds conv1d.weight[o, h*M + m, 0], Mp = the field count rounded up to a power of
 * two (>= 8; columns with m >= M and columns >= H*Mp are zero) — so that 8 consecutive z channels share one h.
 * Same rounding points as oov_cin_outer + oov_tc_linear + oov_cin_pool_dot.
 * Shapes: even M <= 64, H <= 64, O <= 128, even n_hidden / ld_h — oov_cin_layer_supported tells; other shapes take
 * the three-call path. */
int oov_cin_layer_supported(int32_t H, int32_t M, int32_t O, int32_t n_hidden, int64_t ld_h);
int oov_cin_layer(const void* xi, int64_t ld_xi, int32_t H, const void* x0, int64_t ld_x0, int32_t M, int32_t Mp,
                  int64_t B, int32_t D, const void* W, int64_t ldw, const float* bias, int32_t O,
                  void* hid_out, int64_t ld_h, int32_t n_hidden,
                  int32_t pool_lo, int32_t pool_n, const float* pool_w, float* out_acc, void* stream);

/* ------------------------------------------------------------------------------------
 * Sampled-candidate evaluation (eval side) — replaces trainer.py:547-564 /
 * inductive/evaluator.py:116-133 (neg_sample_batch_eval: model.predict on (user, item) pairs, scatter into a
 * [users, N] matrix of -inf, torch.topk) without the [users, N] matrix.
 * rowptr [U + 1] / cols [n_pairs]: the pairs as a CSR by batch user, item ids ascending per row (oov_pairs_to_csr);
 * user_e [U, D]; item_e [n_pairs, D]: the embedding of cols[j] in row j (fp32 or bf16 each).
 * normalize != 0: both rows are L2-normalised first (DirectAU.predict, directau.py:75-78,174-181); 0: plain dot product
 * (BPR.predict, bpr.py:146-149).
 * out [U, k]: (score desc, id asc) among the row's distinct candidates with seg_lo <= id < seg_hi; missing slots are
 * (-inf, -1).  keys: caller workspace of n_pairs 8-byte words; compute_keys = 0 reuses the keys of a previous call
 * (another segment of the same batch).
 * ------------------------------------------------------------------------------------ */
int oov_pair_topk(const void* user_e, int32_t u_dtype, const void* item_e, int32_t i_dtype, int32_t D,
                  const int32_t* rowptr, const int32_t* cols, int64_t U, int64_t n_pairs, int32_t normalize, int32_t k, int64_t seg_lo, int64_t seg_hi,
                  unsigned long long* keys, int32_t compute_keys, float* out_scores, int64_t* out_idx, void* stream);

/* inductive_mapper=random (inductive/random_mapper.py:70-130): new_id = id (id < n_old) or
 * n_old + hash(id - n_old) % n_buckets.  fn: 0 mod, 1 fast, 2 3round, 3 64bit. */
int oov_map_ids(const int64_t* ids, int64_t n, int64_t n_old, int64_t n_buckets, int32_t fn,
                int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OOV_B200_H_ */
