// Backward of the OOV assemble (training side; reference trainer/trainer.py:1748-1837 _train_oov ->
// model.calculate_loss -> bpr.py:48-125 get_user_embedding / get_item_embedding under autograd).
//
// The forward pass is the eval-path assemble kernel run with the training-mode id rule (rows->prime_pad).  Its gradient
// with respect to the tables is a row scatter-add:
//   in-vocab rows (id < n_old):      d table[id]            += g_i      (nn.Embedding backward)       scatter_add_rows
//   mapper / slsh OOV rows:          d buckets[bucket_i]    += g_i      (bpr.py:71, single_lsh:108)   scatter_add_rows
//   lsh OOV rows, out = (H W) / |H|: d buckets[b]           += H_ib g_i / |H_i|  (lsh_embedder.py:156-158)  lsh_backward
// All accumulation is fp32.
#include "common.cuh"

namespace oov {

// one warp per row; red.global.add.f32 on the destination row
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ idx, int64_t idx_stride, int64_t n,
                        int64_t idx_offset, int64_t rows, int D, float* __restrict__ dtable) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t r = idx[i * idx_stride] + idx_offset;
        if (r < 0 || r >= rows) continue;
        for (int d = lane; d < D; d += 32) atomicAdd(dtable + r * D + d, g[i * ldg + d]);
    }
}

// inv[i] = 1 / popcount(bits_i) for OOV rows (inf for an all-zero hash, like the reference's division), 0 for in-vocab rows
__global__ void lsh_inv_count_kernel(const uint32_t* __restrict__ bits, int words, int B, const int64_t* __restrict__ ids,
                                     int64_t ids_stride, int64_t n, int64_t n_old, float* __restrict__ inv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (ids[i * ids_stride] < n_old) { inv[i] = 0.f; continue; }
        int c = 0;
        for (int w = 0; w < words; ++w) {
            uint32_t v = bits[i * words + w];
            if (w == words - 1 && (B & 31)) v &= (1u << (B & 31)) - 1u;
            c += __popc(v);
        }
        inv[i] = 1.f / (float)c;
    }
}

// A block owns 32 buckets (one word column of the multi-hot matrix) x DT columns of the gradient and one slice of the
// ids; a thread keeps `DT / 8` accumulators for (bucket = tid / 8, columns (tid % 8) + 8 j).  The slice results go to
// dW with one atomic per element and block.  g rows and bit words are staged through shared memory 64 ids at a time.
constexpr int LB_IDS = 64;
template <int DT>
__global__ void __launch_bounds__(256)
lsh_backward_kernel(const uint32_t* __restrict__ bits, int words, int B, const float* __restrict__ g, int64_t ldg,
                    const float* __restrict__ inv, int64_t n, int D, int64_t ids_per_slice, float* __restrict__ dW) {
    __shared__ float gs[LB_IDS][DT + 1];
    __shared__ uint32_t ws[LB_IDS];
    __shared__ float is[LB_IDS];
    const int word = blockIdx.x, d0 = blockIdx.y * DT;
    const int64_t lo = (int64_t)blockIdx.z * ids_per_slice;
    const int64_t hi = lo + ids_per_slice < n ? lo + ids_per_slice : n;
    const int b = threadIdx.x >> 3, dl = threadIdx.x & 7;
    float acc[DT / 8];
#pragma unroll
    for (int j = 0; j < DT / 8; ++j) acc[j] = 0.f;
    for (int64_t i0 = lo; i0 < hi; i0 += LB_IDS) {
        const int cnt = (int)(hi - i0 < LB_IDS ? hi - i0 : LB_IDS);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * DT; e += 256) {
            const int r = e / DT, c = e - r * DT;
            gs[r][c] = (d0 + c < D) ? g[(i0 + r) * ldg + d0 + c] : 0.f;
        }
        if (threadIdx.x < cnt) {
            ws[threadIdx.x] = bits[(i0 + threadIdx.x) * words + word];
            is[threadIdx.x] = inv[i0 + threadIdx.x];
        }
        __syncthreads();
        for (int r = 0; r < cnt; ++r) {
            const float s = is[r];
            if (s == 0.f) continue;                                   // in-vocab row
            // (H^T (g / |H|))[b]: a set bit adds g * s; an all-zero hash (s = inf) gives 0 * inf = NaN on EVERY bucket,
            // exactly what autograd produces for lsh_embedder.py:157
            // (dense multiply, no skipping of unset bits: a non-finite g / |H| row reaches every bucket as 0 * x = NaN there too)
            const float h = ((ws[r] >> b) & 1u) ? 1.f : 0.f;
#pragma unroll
            for (int j = 0; j < DT / 8; ++j) acc[j] += h * (gs[r][dl + 8 * j] * s);
        }
    }
    const int bucket = word * 32 + b;
    if (bucket < B) {
#pragma unroll
        for (int j = 0; j < DT / 8; ++j) {
            const int d = d0 + dl + 8 * j;
            if (d < D && acc[j] != 0.f) atomicAdd(dW + (int64_t)bucket * D + d, acc[j]);
        }
    }
}

}  // namespace oov

using namespace oov;

extern "C" {

int oov_scatter_add_rows(const float* g, int64_t ldg, const int64_t* idx, int64_t idx_stride, int64_t n, int64_t idx_offset,
                         int64_t rows, int32_t D, float* dtable, void* stream) {
    OOV_REQUIRE(n >= 0 && rows >= 0 && D > 0 && ldg >= D && idx_stride >= 1, OOV_ERR_ARG, "oov_scatter_add_rows: bad shape");
    if (n == 0 || rows == 0) return OOV_OK;
    OOV_REQUIRE(g && idx && dtable, OOV_ERR_ARG, "oov_scatter_add_rows: NULL pointer");
    int64_t blocks = cdiv(n, 8);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    scatter_add_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, idx, idx_stride, n, idx_offset, rows, D, dtable);
    OOV_LAUNCH_CHECK("scatter_add_rows_kernel");
    return OOV_OK;
}

size_t oov_lsh_embed_backward_workspace(int64_t n) { return n > 0 ? align_up((size_t)n * 4, 256) : 0; }

int oov_lsh_embed_backward(const uint32_t* bits, int32_t B, const float* g, int64_t ldg, const int64_t* ids, int64_t ids_stride,
                           int64_t n, int64_t n_old, int32_t D, float* dW, void* workspace, size_t workspace_bytes, void* stream) {
    OOV_REQUIRE(n >= 0 && B > 0 && D > 0 && ldg >= D && ids_stride >= 1, OOV_ERR_ARG, "oov_lsh_embed_backward: bad shape");
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(bits && g && ids && dW, OOV_ERR_ARG, "oov_lsh_embed_backward: NULL pointer");
    OOV_REQUIRE(workspace && workspace_bytes >= (size_t)n * 4, OOV_ERR_WORKSPACE, "oov_lsh_embed_backward: workspace %zu < %zu",
                workspace_bytes, (size_t)n * 4);
    cudaStream_t st = (cudaStream_t)stream;
    float* inv = reinterpret_cast<float*>(workspace);
    const int words = (B + 31) / 32;
    lsh_inv_count_kernel<<<(unsigned)(cdiv(n, 256) < 1184 ? cdiv(n, 256) : 1184), 256, 0, st>>>(bits, words, B, ids, ids_stride, n, n_old, inv);
    OOV_LAUNCH_CHECK("lsh_inv_count_kernel");
    constexpr int DT = 64;
    const int dy = (int)cdiv(D, DT);
    // enough id slices to fill the machine, at least 256 ids each
    int64_t slices = cdiv((int64_t)num_sms() * 2, (int64_t)words * dy);
    if (slices > cdiv(n, 256)) slices = cdiv(n, 256);
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    const int64_t per = cdiv(cdiv(n, slices), LB_IDS) * LB_IDS;
    slices = cdiv(n, per);
    lsh_backward_kernel<DT><<<dim3((unsigned)words, (unsigned)dy, (unsigned)slices), 256, 0, st>>>(bits, words, B, g, ldg, inv, n, D, per, dW);
    OOV_LAUNCH_CHECK("lsh_backward_kernel");
    return OOV_OK;
}

}  // extern "C"
