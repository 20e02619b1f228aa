// xDeepFM compressed interaction network (reference model/context_aware_recommender/xdeepfm.py:134-190), the part
// around the tensor-core linear:
//   z^k[b, h*M + m, d] = X^{k-1}[b, h, d] * X^0[b, m, d]      (einsum "bhd,bmd->bhmd" + view)       -> cin_outer_kernel
//   X^k = ReLU(conv1d_k(z^k))  (kernel size 1 = a linear over the channel axis, per (b, d))          -> oov_tc_linear
//   p = sum_d direct_connect[b, :, d];  cin_linear(p)                                               -> cin_pool_dot_kernel
// Rows of every matrix here are (b, d) pairs, channels run along the row ("d-major"): z^k is [B*D, H_{k-1}*M] bf16, the
// layout the linear takes as its A operand, and its output [B*D, H_k] is directly the next layer's X^k in the same layout.
#include "common.cuh"

namespace oov {

// element (b, d, c) of an operand sits at base[b * sb + d * sd + c * sc]  (X^0 as the gather wrote it, [B, M, D]:
// sb = M*D, sd = 1, sc = D; a linear's output [B*D, ld]: sb = D*ld, sd = ld, sc = 1)
struct CinView { const __nv_bfloat16* p; int64_t sb, sd, sc; };

// 8 channels (16 bytes of z) per thread and step; padding channels [H*M, ldz) are written as zeros
__global__ void __launch_bounds__(256)
cin_outer_kernel(CinView xi, int H, CinView x0, int M, int64_t B, int D, __nv_bfloat16* __restrict__ z, int64_t ldz) {
    const int c8n = (int)(ldz >> 3);
    const int64_t total = B * (int64_t)D * c8n;
    const int C = H * M;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / c8n;
        const int c0 = (int)(t - r * c8n) << 3;
        const int64_t b = r / D;
        const int d = (int)(r - b * D);
        const __nv_bfloat16* pi = xi.p + b * xi.sb + d * xi.sd;
        const __nv_bfloat16* p0 = x0.p + b * x0.sb + d * x0.sd;
        int h = c0 / M, m = c0 - h * M;
        float a = (c0 < C) ? __bfloat162float(pi[h * xi.sc]) : 0.f;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                v[e] = (c0 + j + e < C) ? a * __bfloat162float(p0[m * x0.sc]) : 0.f;
                if (++m == M) {
                    m = 0;
                    ++h;
                    a = (h < H) ? __bfloat162float(pi[h * xi.sc]) : 0.f;
                }
            }
            __nv_bfloat162 pk = __floats2bfloat162_rn(v[0], v[1]);
            w[j >> 1] = *reinterpret_cast<uint32_t*>(&pk);
        }
        *reinterpret_cast<uint4*>(z + r * ldz + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// acc[b] (+)= bias + sum_{d < D, c < ncols} y[(b*D + d) * ldy + col0 + c] * w[c]: sum pooling over the embedding axis
// (xdeepfm.py:188-189) folded with this layer's slice of cin_linear (xdeepfm.py:198).  One warp per batch row.
__global__ void __launch_bounds__(256)
cin_pool_dot_kernel(const __nv_bfloat16* __restrict__ y, int64_t ldy, int col0, int ncols, int64_t B, int D,
                    const float* __restrict__ w, float bias, int accumulate, float* __restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += nwarps) {
        float s = 0.f;
        for (int c = lane; c < ncols; c += 32) {
            float p = 0.f;
            for (int d = 0; d < D; ++d) p += __bfloat162float(y[(b * D + d) * ldy + col0 + c]);
            s = fmaf(p, __ldg(w + c), s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) acc[b] = (accumulate ? acc[b] : 0.f) + bias + s;
    }
}

static unsigned cin_grid(int64_t work_items) {
    int64_t blocks = cdiv(work_items, 256);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace oov

using namespace oov;

extern "C" {

int oov_cin_outer(const void* xi, int64_t xi_sb, int64_t xi_sd, int64_t xi_sc, int32_t H,
                  const void* x0, int64_t x0_sb, int64_t x0_sd, int64_t x0_sc, int32_t M,
                  int64_t B, int32_t D, void* z, int64_t ldz, void* stream) {
    OOV_REQUIRE(B >= 0 && D > 0 && H > 0 && M > 0, OOV_ERR_ARG, "oov_cin_outer: bad shape B=%lld D=%d H=%d M=%d", (long long)B, D, H, M);
    OOV_REQUIRE(ldz % 8 == 0 && ldz >= (int64_t)H * M, OOV_ERR_ALIGN, "oov_cin_outer: ldz=%lld must be a multiple of 8 >= H*M", (long long)ldz);
    if (B == 0) return OOV_OK;
    OOV_REQUIRE(xi && x0 && z && aligned(z, 16), OOV_ERR_ARG, "oov_cin_outer: NULL / misaligned pointer");
    const CinView vi{reinterpret_cast<const __nv_bfloat16*>(xi), xi_sb, xi_sd, xi_sc};
    const CinView v0{reinterpret_cast<const __nv_bfloat16*>(x0), x0_sb, x0_sd, x0_sc};
    cin_outer_kernel<<<cin_grid(B * D * (ldz / 8)), 256, 0, (cudaStream_t)stream>>>(vi, H, v0, M, B, D,
                                                                                  reinterpret_cast<__nv_bfloat16*>(z), ldz);
    OOV_LAUNCH_CHECK("cin_outer_kernel");
    return OOV_OK;
}

int oov_cin_pool_dot(const void* y, int64_t ldy, int32_t col0, int32_t ncols, int64_t B, int32_t D,
                     const float* w, float bias, int32_t accumulate, float* acc, void* stream) {
    OOV_REQUIRE(B >= 0 && D > 0 && col0 >= 0 && ncols > 0 && ldy >= col0 + ncols, OOV_ERR_ARG, "oov_cin_pool_dot: bad shape");
    if (B == 0) return OOV_OK;
    OOV_REQUIRE(y && w && acc, OOV_ERR_ARG, "oov_cin_pool_dot: NULL pointer");
    cin_pool_dot_kernel<<<cin_grid(B * 32), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(y), ldy, col0,
                                                                          ncols, B, D, w, bias, accumulate, acc);
    OOV_LAUNCH_CHECK("cin_pool_dot_kernel");
    return OOV_OK;
}

}  // extern "C"
