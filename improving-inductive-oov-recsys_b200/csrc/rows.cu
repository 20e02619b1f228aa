// HBM-bound row movers: column mean, constant (mean/zero) embed, row gather, context token
// gather with OOV overwrite, first-order sum, random-mapper integer hashes.
//
// Restates inductive/mean_embedder.py:42-87, zero_embedder.py:36-60, the nn.Embedding gathers of
// model/general_recommender/bpr.py:48-125, model/abstract_recommender.py:794-842 +
// model/layers.py:150-153,1634-1693, and inductive/random_mapper.py:70-130 of the reference.
// All of these are bandwidth-bound byte movers: 128-bit accesses, grid = multiple of the SM count.
#include "common.cuh"

namespace oov {

int check_rows_public(const oov_rows* r, const char* who);

// ---------------------------------------------------------------- column mean (two deterministic passes)
constexpr int CM_THREADS = 256;
constexpr int CM_ROWS_PER_BLOCK = 4096;

// partial[blk, d] = sum over the block's row slab (fp32, fixed order): thread t owns column d = t % Dp,
// row phase t / Dp; phases are combined through smem in a fixed order.
__global__ void __launch_bounds__(CM_THREADS)
col_sum_partial(const void* __restrict__ table, int dtype, int64_t rows, int D, float* __restrict__ partial) {
    extern __shared__ float red[];                 // [phases][D]
    const int phases = max(1, CM_THREADS / D);
    const int d = threadIdx.x % D, ph = threadIdx.x / D;
    const int64_t r0 = (int64_t)blockIdx.x * CM_ROWS_PER_BLOCK;
    const int64_t r1 = min(rows, r0 + CM_ROWS_PER_BLOCK);
    float acc = 0.f;
    if (threadIdx.x < phases * D) {
        for (int64_t r = r0 + ph; r < r1; r += phases) acc += load_elem(table, dtype, r * D + d);
        red[ph * D + d] = acc;
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float s = 0.f;
        for (int p = 0; p < phases; ++p) s += red[p * D + threadIdx.x];
        partial[(int64_t)blockIdx.x * D + threadIdx.x] = s;
    }
}
// D > CM_THREADS: one thread handles several columns
__global__ void __launch_bounds__(CM_THREADS)
col_sum_partial_wide(const void* __restrict__ table, int dtype, int64_t rows, int D, float* __restrict__ partial) {
    const int64_t r0 = (int64_t)blockIdx.x * CM_ROWS_PER_BLOCK;
    const int64_t r1 = min(rows, r0 + CM_ROWS_PER_BLOCK);
    for (int d = threadIdx.x; d < D; d += CM_THREADS) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc += load_elem(table, dtype, r * D + d);
        partial[(int64_t)blockIdx.x * D + d] = acc;
    }
}
__global__ void col_mean_final(const float* __restrict__ partial, int nblk, int D, int64_t rows, float* __restrict__ mean) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * D + d];
    mean[d] = s / (float)rows;
}

// ---------------------------------------------------------------- constant embed + assemble
// 16-lane groups, one row per group
__global__ void __launch_bounds__(256)
const_embed_kernel(const float* __restrict__ vec, const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n,
                   int64_t n_old, const void* __restrict__ iv_table, int iv_dtype,
                   void* __restrict__ out, int out_dtype, int64_t out_stride, int D) {
    const int sub = threadIdx.x & 15;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2, isz = iv_dtype == OOV_F32 ? 4 : 2;
    for (int64_t r = g0; r < n; r += ng) {
        const int64_t id = ids[r * ids_stride];
        char* orow = reinterpret_cast<char*>(out) + (size_t)r * out_stride * osz;
        if (id < n_old) {
            if (iv_table != nullptr && id >= 0)
                copy_row(reinterpret_cast<const char*>(iv_table) + (size_t)id * D * isz, iv_dtype, orow, out_dtype, D, sub, 16);
        } else {
            for (int d = sub; d < D; d += 16) store_elem(orow, out_dtype, d, vec ? vec[d] : 0.f);
        }
    }
}

// OOV-only id lists (n_old <= 0: the OOV half of an assembled table) with 16-byte-multiple output rows: every lane keeps
// its 16-byte piece of the converted constant in a register and the kernel is a pure coalesced store stream (the kernel
// above reads the id first and stores 4-byte elements: 23-42 % of the HBM rate).
__global__ void __launch_bounds__(256)
const_fill_rows_kernel(const float* __restrict__ vec, int64_t n, void* __restrict__ out, int out_dtype, int D) {
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2;
    const int row_bytes = D * (int)osz, lpr = row_bytes >> 4;               // lanes per row
    __shared__ uint4 piece_s[64];
    if (threadIdx.x < lpr) {
        uint32_t w[4];
        for (int q = 0; q < 4; ++q) {
            if (out_dtype == OOV_F32) {
                w[q] = __float_as_uint(vec ? vec[threadIdx.x * 4 + q] : 0.f);
            } else {
                const float a = vec ? vec[threadIdx.x * 8 + 2 * q] : 0.f, b = vec ? vec[threadIdx.x * 8 + 2 * q + 1] : 0.f;
                __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                w[q] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        piece_s[threadIdx.x] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    const int64_t total = n * lpr;                                          // 16-byte pieces of the whole output
    uint4* o = reinterpret_cast<uint4*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = piece_s[(int)(i % lpr)];
}

// ---------------------------------------------------------------- plain gather
__global__ void __launch_bounds__(256)
gather_rows_kernel(const void* __restrict__ table, int dtype, int64_t table_rows, int D,
                   const int64_t* __restrict__ idx, int64_t idx_stride, int64_t n, int64_t idx_offset,
                   void* __restrict__ out, int out_dtype, int64_t out_stride) {
    const int sub = threadIdx.x & 15;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2, isz = dtype == OOV_F32 ? 4 : 2;
    for (int64_t r = g0; r < n; r += ng) {
        const int64_t id = idx[r * idx_stride] + idx_offset;
        if (id < 0 || id >= table_rows) continue;            // nn.Embedding would raise; never write garbage
        copy_row(reinterpret_cast<const char*>(table) + (size_t)id * D * isz, dtype,
                 reinterpret_cast<char*>(out) + (size_t)r * out_stride * osz, out_dtype, D, sub, 16);
    }
}

// fp32 table -> bf16 rows, D % 4 == 0, 16-byte aligned rows (the in-vocab half of every assembled bf16 item table):
// four rows per 16-lane group and iteration, all ids first, then all row loads, then the stores — the plain kernel
// walks one row at a time behind a dependent id load and reaches a fifth of the HBM rate.
__global__ void __launch_bounds__(256)
gather_rows_f32_bf16_kernel(const float* __restrict__ table, int64_t table_rows, int D,
                            const int64_t* __restrict__ idx, int64_t idx_stride, int64_t n, int64_t idx_offset,
                            __nv_bfloat16* __restrict__ out, int64_t out_stride) {
    const int sub = threadIdx.x & 15;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const int d4 = D >> 2;
    for (int64_t r0 = g0; r0 < n; r0 += 4 * ng) {
        int64_t id[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r0 + j * ng;
            id[j] = r < n ? idx[r * idx_stride] + idx_offset : -1;
        }
        for (int i = sub; i < d4; i += 16) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (id[j] >= 0 && id[j] < table_rows) v[j] = __ldg(reinterpret_cast<const float4*>(table + (size_t)id[j] * D) + i);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (id[j] < 0 || id[j] >= table_rows) continue;
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[j].x, v[j].y), hi = __floats2bfloat162_rn(v[j].z, v[j].w);
                reinterpret_cast<uint2*>(out + (size_t)(r0 + j * ng) * out_stride)[i] =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

// ---------------------------------------------------------------- context token gather (+ OOV overwrite)
// one LPG-lane group per (row, field); LPG = 4 for D <= 16, 16 otherwise
template <int LPG>
__global__ void __launch_bounds__(256)
token_gather_kernel(const int64_t* __restrict__ tokens, int64_t Bn, int fields, const int64_t* __restrict__ offsets,
                    const void* __restrict__ table, int dtype, int64_t table_rows, int D,
                    int64_t n_users, int64_t n_items, int uid_idx, int iid_idx,
                    const float* __restrict__ user_const, const float* __restrict__ item_const,
                    void* __restrict__ out, int out_dtype) {
    const int sub = threadIdx.x % LPG;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) / LPG;
    const int64_t total = Bn * (int64_t)fields;
    const size_t osz = out_dtype == OOV_F32 ? 4 : 2, isz = dtype == OOV_F32 ? 4 : 2;
    for (int64_t e = g0; e < total; e += ng) {
        const int f = (int)(e % fields);
        const int64_t id = tokens[e];
        char* orow = reinterpret_cast<char*>(out) + (size_t)e * D * osz;
        const float* cvec = nullptr;
        bool oov = false;
        if (f == uid_idx && id >= n_users) { oov = true; cvec = user_const; }
        else if (f == iid_idx && id >= n_items) { oov = true; cvec = item_const; }
        if (oov) {
            // abstract_recommender.py:818-836: looked up as id 0, then overwritten by the embedder output;
            // only the overwrite is observable, so the lookup is skipped.
            if (cvec != nullptr)
                for (int d = sub; d < D; d += LPG) store_elem(orow, out_dtype, d, cvec[d]);
            continue;
        }
        const int64_t trow = id + offsets[f];
        if (trow < 0 || trow >= table_rows) continue;
        copy_row(reinterpret_cast<const char*>(table) + (size_t)trow * D * isz, dtype, orow, out_dtype, D, sub, LPG);
    }
}

// Same-dtype rows whose byte length is a multiple of PB (16 / 8 / 4): four cells per lane group and iteration, all token
// ids first, then all table pieces, then the stores.  The kernel above walks token -> offset -> row -> store one cell at
// a time (two dependent memory latencies per cell) and reached 26-58 % of the HBM rate of its algorithmic bytes.
template <int PB> struct PieceT;
template <> struct PieceT<16> { using T = uint4; };
template <> struct PieceT<8> { using T = uint2; };
template <> struct PieceT<4> { using T = uint32_t; };

template <int LPG, int PB, int MAXK>
__global__ void __launch_bounds__(256)
token_gather_vec_kernel(const int64_t* __restrict__ tokens, int64_t Bn, int fields, const int64_t* __restrict__ offsets,
                        const void* __restrict__ table, int dtype, int64_t table_rows, int D,
                        int64_t n_users, int64_t n_items, int uid_idx, int iid_idx,
                        const float* __restrict__ user_const, const float* __restrict__ item_const,
                        void* __restrict__ out) {
    using P = typename PieceT<PB>::T;
    const int sub = threadIdx.x % LPG;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) / LPG;
    const int64_t total = Bn * (int64_t)fields;
    const size_t esz = dtype == OOV_F32 ? 4 : 2;
    const int row_bytes = D * (int)esz, npieces = row_bytes / PB;
    const bool small = total < (1ll << 31);
    for (int64_t e0 = g0; e0 < total; e0 += 4 * ng) {
        int64_t id[4];
        int f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t e = e0 + j * ng;
            id[j] = e < total ? __ldg(tokens + e) : -1;
            f[j] = e < total ? (small ? (int)((uint32_t)e % (uint32_t)fields) : (int)(e % fields)) : 0;
        }
        const char* src[4];
        const float* cvec[4];
        bool oov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            src[j] = nullptr; cvec[j] = nullptr; oov[j] = false;
            if (e0 + j * ng >= total) continue;
            if (f[j] == uid_idx && id[j] >= n_users) { oov[j] = true; cvec[j] = user_const; }
            else if (f[j] == iid_idx && id[j] >= n_items) { oov[j] = true; cvec[j] = item_const; }
            else {
                const int64_t trow = id[j] + __ldg(offsets + f[j]);
                if (trow >= 0 && trow < table_rows) src[j] = reinterpret_cast<const char*>(table) + (size_t)trow * row_bytes;
            }
        }
        P v[4][MAXK];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int kk = 0; kk < MAXK; ++kk) {
                const int pc = sub + kk * LPG;
                if (src[j] != nullptr && pc < npieces) v[j][kk] = __ldg(reinterpret_cast<const P*>(src[j]) + pc);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t e = e0 + j * ng;
            if (e >= total) continue;
            char* orow = reinterpret_cast<char*>(out) + (size_t)e * row_bytes;
            if (oov[j]) {
                // abstract_recommender.py:818-836: looked up as id 0, then overwritten by the embedder output
                if (cvec[j] != nullptr)
                    for (int d = sub; d < D; d += LPG) store_elem(orow, dtype, d, cvec[j][d]);
                continue;
            }
            if (src[j] == nullptr) continue;
#pragma unroll
            for (int kk = 0; kk < MAXK; ++kk) {
                const int pc = sub + kk * LPG;
                if (pc < npieces) reinterpret_cast<P*>(orow)[pc] = v[j][kk];
            }
        }
    }
}

// first-order: one thread per row, D = 1
__global__ void __launch_bounds__(256)
first_order_sum_kernel(const int64_t* __restrict__ tokens, int64_t Bn, int fields, const int64_t* __restrict__ offsets,
                       const float* __restrict__ table1, int64_t table_rows,
                       int64_t n_users, int64_t n_items, int uid_idx, int iid_idx,
                       const float* __restrict__ oov_user_val, const float* __restrict__ oov_item_val,
                       float* __restrict__ out) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < Bn; b += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int f = 0; f < fields; ++f) {          // torch.sum over dim=1 in field order (layers.py:1692)
            const int64_t id = tokens[b * fields + f];
            float v;
            if (f == uid_idx && id >= n_users) v = oov_user_val ? oov_user_val[b] : 0.f;
            else if (f == iid_idx && id >= n_items) v = oov_item_val ? oov_item_val[b] : 0.f;
            else {
                const int64_t trow = id + offsets[f];
                v = (trow >= 0 && trow < table_rows) ? __ldg(table1 + trow) : 0.f;
            }
            s += v;
        }
        out[b] = s;
    }
}

// ---------------------------------------------------------------- random mapper hashes
__device__ __forceinline__ int64_t py_mod(int64_t x, int64_t m) { int64_t r = x % m; return (r != 0 && ((r < 0) != (m < 0))) ? r + m : r; }
__device__ __forceinline__ int64_t mul_wrap(int64_t a, uint64_t c) { return (int64_t)((uint64_t)a * c); }

__global__ void map_ids_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t n_old, int64_t nb, int fn, int64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = ids[i];
        if (id < n_old) { out[i] = id; continue; }
        int64_t x = id - n_old, h;
        if (fn == 0) h = py_mod(x, nb);
        else if (fn == 1) {      // int64 tensors: wrapping multiply, arithmetic >> (random_mapper.py:70-76)
            x ^= x >> 16; x = mul_wrap(x, 0x21f0aaadull); x ^= x >> 15; x = mul_wrap(x, 0xd35a2d97ull); x ^= x >> 15;
            h = py_mod(x, nb);
        } else if (fn == 2) {
            x ^= x >> 17; x = mul_wrap(x, 0xed5ad4bbull); x ^= x >> 11; x = mul_wrap(x, 0xac4c1b51ull);
            x ^= x >> 15; x = mul_wrap(x, 0x31848babull); x ^= x >> 14;
            h = py_mod(x, nb);
        } else {                 // numpy uint64 (random_mapper.py:95-102)
            uint64_t u = (uint64_t)x;
            u = (u ^ (u >> 30)) * 0xb9e5e41c6d4758bfull;
            u = (u ^ (u >> 27)) * 0xeb113113bb49d094ull;
            u = u ^ (u >> 31);
            h = (int64_t)(u % (uint64_t)nb);
        }
        out[i] = h + n_old;
    }
}

// Grid of a grid-stride kernel: enough blocks for the work, capped at ONE resident wave (blocks per SM from the
// occupancy calculator).  A fixed cap of 8 blocks per SM left kernels that fit 4-5 blocks per SM with a partial second
// wave that ran on a fraction of the machine.
template <typename K>
static unsigned grid_for(K kernel, int64_t work_items, int per_block) {
    static int per_sm_dev[64] = {0};             // one instantiation (and one cache slot per device) per kernel type
    static const void* cached_for[64] = {nullptr};
    const int dev = cur_device();
    if (cached_for[dev] != reinterpret_cast<const void*>(kernel)) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, 256, 0) != cudaSuccess || v < 1) v = 4;
        per_sm_dev[dev] = v;
        cached_for[dev] = reinterpret_cast<const void*>(kernel);
    }
    const int per_sm = per_sm_dev[dev];
    int64_t b = cdiv(work_items, per_block);
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// DCN-V2 cross layer tail (dcnv2.py:120-144): x_{l+1} = x_0 * (W_l x_l + b_l) + x_l, elementwise on bf16 rows; the
// matrix-vector part t = W_l x_l + b_l comes from the tensor-core linear.  8 elements (16 bytes) per thread and step.
__global__ void __launch_bounds__(256)
cross_update_kernel(const uint4* __restrict__ x0, const uint4* __restrict__ t, const uint4* __restrict__ xl, int64_t n8,
                    uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(x0 + i), b = __ldg(t + i), c = __ldg(xl + i);
        uint4 o;
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
        const uint32_t* pc = reinterpret_cast<const uint32_t*>(&c);
        uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pa[j]));
            const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pb[j]));
            const float2 fc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pc[j]));
            __nv_bfloat162 h = __floats2bfloat162_rn(fmaf(fa.x, fb.x, fc.x), fmaf(fa.y, fb.y, fc.y));
            po[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        out[i] = o;
    }
}

}  // namespace oov

using namespace oov;

extern "C" {

size_t oov_col_mean_workspace(int64_t rows, int32_t D) {
    if (rows <= 0 || D <= 0) return 0;
    return align_up((size_t)cdiv(rows, CM_ROWS_PER_BLOCK) * D * 4, 256);
}

int oov_col_mean(const void* table, int32_t dtype, int64_t rows, int32_t D, float* mean_out,
                 void* workspace, size_t workspace_bytes, void* stream) {
    OOV_REQUIRE(table && mean_out && dtype_ok(dtype) && rows > 0 && D > 0, OOV_ERR_ARG, "oov_col_mean: bad argument");
    const int nblk = (int)cdiv(rows, CM_ROWS_PER_BLOCK);
    OOV_REQUIRE(workspace && workspace_bytes >= (size_t)nblk * D * 4, OOV_ERR_WORKSPACE, "oov_col_mean: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    if (D <= CM_THREADS) {
        const int phases = CM_THREADS / D > 0 ? CM_THREADS / D : 1;
        col_sum_partial<<<nblk, CM_THREADS, (size_t)phases * D * 4, st>>>(table, dtype, rows, D, partial);
    } else {
        col_sum_partial_wide<<<nblk, CM_THREADS, 0, st>>>(table, dtype, rows, D, partial);
    }
    OOV_LAUNCH_CHECK("col_sum_partial");
    col_mean_final<<<(unsigned)cdiv(D, 128), 128, 0, st>>>(partial, nblk, D, rows, mean_out);
    OOV_LAUNCH_CHECK("col_mean_final");
    return OOV_OK;
}

int oov_const_embed(const float* vec, const oov_rows* rows, void* stream) {
    int rc = check_rows_public(rows, "oov_const_embed");
    if (rc) return rc;
    if (rows->n == 0) return OOV_OK;
    {
        const int row_bytes = rows->D * (rows->out_dtype == OOV_F32 ? 4 : 2);
        if (rows->n_old <= 0 && rows->out_stride == rows->D && row_bytes % 16 == 0 && row_bytes <= 1024 && aligned(rows->out, 16)) {
            // ids >= 0 >= n_old: every row is OOV, the ids need not be read
            const_fill_rows_kernel<<<grid_for(const_fill_rows_kernel, rows->n * (row_bytes / 16), 256), 256, 0, (cudaStream_t)stream>>>(
                vec, rows->n, rows->out, rows->out_dtype, rows->D);
            OOV_LAUNCH_CHECK("const_fill_rows_kernel");
            return OOV_OK;
        }
    }
    const_embed_kernel<<<grid_for(const_embed_kernel, rows->n, 16), 256, 0, (cudaStream_t)stream>>>(
        vec, rows->ids, rows->ids_stride, rows->n, rows->n_old, rows->iv_table, rows->iv_dtype, rows->out, rows->out_dtype,
        rows->out_stride, rows->D);
    OOV_LAUNCH_CHECK("const_embed_kernel");
    return OOV_OK;
}

int oov_gather_rows(const void* table, int32_t dtype, int64_t table_rows, int32_t D, const int64_t* idx,
                    int64_t idx_stride, int64_t n, int64_t idx_offset, void* out, int32_t out_dtype, int64_t out_stride,
                    void* stream) {
    OOV_REQUIRE(table && dtype_ok(dtype) && dtype_ok(out_dtype) && table_rows > 0 && D > 0 && n >= 0 && idx_stride >= 1 &&
                    out_stride >= D, OOV_ERR_ARG, "oov_gather_rows: bad argument");
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(idx && out, OOV_ERR_ARG, "oov_gather_rows: NULL pointer");
    if (dtype == OOV_F32 && out_dtype == OOV_BF16 && D % 4 == 0 && out_stride % 4 == 0 && aligned(table, 16) && aligned(out, 8)) {
        gather_rows_f32_bf16_kernel<<<grid_for(gather_rows_f32_bf16_kernel, n, 64), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float*>(table), table_rows, D, idx, idx_stride, n, idx_offset,
            reinterpret_cast<__nv_bfloat16*>(out), out_stride);
        OOV_LAUNCH_CHECK("gather_rows_f32_bf16_kernel");
        return OOV_OK;
    }
    gather_rows_kernel<<<grid_for(gather_rows_kernel, n, 16), 256, 0, (cudaStream_t)stream>>>(table, dtype, table_rows, D, idx, idx_stride, n,
                                                                          idx_offset, out, out_dtype, out_stride);
    OOV_LAUNCH_CHECK("gather_rows_kernel");
    return OOV_OK;
}

int oov_token_gather(const int64_t* tokens, int64_t Bn, int32_t fields, const int64_t* offsets, const void* table,
                     int32_t dtype, int64_t table_rows, int32_t D, int64_t n_users, int64_t n_items, int32_t uid_idx,
                     int32_t iid_idx, const float* user_const, const float* item_const, void* out, int32_t out_dtype,
                     void* stream) {
    OOV_REQUIRE(table && offsets && dtype_ok(dtype) && dtype_ok(out_dtype) && Bn >= 0 && fields > 0 && D > 0 && table_rows > 0,
                OOV_ERR_ARG, "oov_token_gather: bad argument");
    if (Bn == 0) return OOV_OK;
    OOV_REQUIRE(tokens && out, OOV_ERR_ARG, "oov_token_gather: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = Bn * fields;
    // vectorised path: same dtype in and out, rows a multiple of 16 / 8 / 4 bytes, at most MAXK pieces per lane
    if (dtype == out_dtype) {
        const int row_bytes = D * (dtype == OOV_F32 ? 4 : 2);
        const int lpg = D <= 16 ? 4 : 16;
        const int pb = (row_bytes % 16 == 0 && aligned(table, 16) && aligned(out, 16)) ? 16
                       : ((row_bytes % 8 == 0 && aligned(table, 8) && aligned(out, 8)) ? 8 : (row_bytes % 4 == 0 ? 4 : 0));
        if (pb != 0 && row_bytes / pb <= lpg * 4) {
#define OOV_TG(LPG_, PB_, MK_)                                                                                                \
            token_gather_vec_kernel<LPG_, PB_, MK_><<<grid_for(token_gather_vec_kernel<LPG_, PB_, MK_>, total, 4 * (256 / LPG_)), 256, 0, st>>>( \
                tokens, Bn, fields, offsets, table, dtype, table_rows, D, n_users, n_items, uid_idx, iid_idx, user_const,     \
                item_const, out)
#define OOV_TG2(LPG_, PB_) do { if (row_bytes / PB_ <= LPG_) OOV_TG(LPG_, PB_, 1); else OOV_TG(LPG_, PB_, 4); } while (0)
            if (lpg == 4) { if (pb == 16) OOV_TG2(4, 16); else if (pb == 8) OOV_TG2(4, 8); else OOV_TG2(4, 4); }
            else { if (pb == 16) OOV_TG2(16, 16); else if (pb == 8) OOV_TG2(16, 8); else OOV_TG2(16, 4); }
#undef OOV_TG2
#undef OOV_TG
            OOV_LAUNCH_CHECK("token_gather_vec_kernel");
            return OOV_OK;
        }
    }
    if (D <= 16)
        token_gather_kernel<4><<<grid_for(token_gather_kernel<4>, total, 64), 256, 0, st>>>(tokens, Bn, fields, offsets, table, dtype, table_rows, D,
                                                                    n_users, n_items, uid_idx, iid_idx, user_const,
                                                                    item_const, out, out_dtype);
    else
        token_gather_kernel<16><<<grid_for(token_gather_kernel<16>, total, 16), 256, 0, st>>>(tokens, Bn, fields, offsets, table, dtype, table_rows, D,
                                                                     n_users, n_items, uid_idx, iid_idx, user_const,
                                                                     item_const, out, out_dtype);
    OOV_LAUNCH_CHECK("token_gather_kernel");
    return OOV_OK;
}

int oov_first_order_sum(const int64_t* tokens, int64_t Bn, int32_t fields, const int64_t* offsets, const float* table1,
                        int64_t table_rows, int64_t n_users, int64_t n_items, int32_t uid_idx, int32_t iid_idx,
                        const float* oov_user_val, const float* oov_item_val, float* out, void* stream) {
    OOV_REQUIRE(table1 && offsets && Bn >= 0 && fields > 0 && table_rows > 0, OOV_ERR_ARG, "oov_first_order_sum: bad argument");
    if (Bn == 0) return OOV_OK;
    OOV_REQUIRE(tokens && out, OOV_ERR_ARG, "oov_first_order_sum: NULL pointer");
    first_order_sum_kernel<<<grid_for(first_order_sum_kernel, Bn, 256), 256, 0, (cudaStream_t)stream>>>(
        tokens, Bn, fields, offsets, table1, table_rows, n_users, n_items, uid_idx, iid_idx, oov_user_val, oov_item_val, out);
    OOV_LAUNCH_CHECK("first_order_sum_kernel");
    return OOV_OK;
}

int oov_map_ids(const int64_t* ids, int64_t n, int64_t n_old, int64_t n_buckets, int32_t fn, int64_t* out, void* stream) {
    OOV_REQUIRE(n >= 0 && n_buckets > 0 && fn >= 0 && fn <= 3, OOV_ERR_ARG, "oov_map_ids: bad argument");
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(ids && out, OOV_ERR_ARG, "oov_map_ids: NULL pointer");
    map_ids_kernel<<<grid_for(map_ids_kernel, n, 256), 256, 0, (cudaStream_t)stream>>>(ids, n, n_old, n_buckets, fn, out);
    OOV_LAUNCH_CHECK("map_ids_kernel");
    return OOV_OK;
}

int oov_cross_update(const void* x0, const void* t, const void* xl, int64_t n_elems, void* out, void* stream) {
    OOV_REQUIRE(n_elems >= 0 && n_elems % 8 == 0, OOV_ERR_ARG, "oov_cross_update: n_elems=%lld must be a multiple of 8", (long long)n_elems);
    if (n_elems == 0) return OOV_OK;
    OOV_REQUIRE(x0 && t && xl && out, OOV_ERR_ARG, "oov_cross_update: NULL pointer");
    OOV_REQUIRE(aligned(x0, 16) && aligned(t, 16) && aligned(xl, 16) && aligned(out, 16), OOV_ERR_ALIGN,
                "oov_cross_update: operands must be 16-byte aligned");
    const int64_t n8 = n_elems / 8;
    cross_update_kernel<<<grid_for(cross_update_kernel, n8, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x0), reinterpret_cast<const uint4*>(t), reinterpret_cast<const uint4*>(xl), n8,
        reinterpret_cast<uint4*>(out));
    OOV_LAUNCH_CHECK("cross_update_kernel");
    return OOV_OK;
}

}  // extern "C"
