// One xDeepFM CIN layer in ONE tcgen05 kernel (reference model/context_aware_recommender/xdeepfm.py:157-189):
//   z[(b, d), h*M + m] = X^{k-1}[b, h, d] * X^0[b, m, d]     — never written to memory
//   Y = ReLU(z W^T + bias)                                    — W = conv1d_k.weight[:, :, 0]  [O, H*M]
//   hidden part  Y[:, :n_hidden]     -> hid_out (the next layer's X^k, rows = (b, d) pairs)
//   direct part  Y[:, lo : lo + n]   -> out_acc[b] += sum_d sum_c Y * pool_w[c]   (sum pooling + this layer's slice of cin_linear)
// The unfused path (cin.cu + oov_tc_linear) writes and re-reads z through HBM: 2 x 1.7 GB per layer at B = 65536, D = 10,
// H*M = 1300 — the layer is bound by that traffic (5.7 ms for the three layers).  Here z only exists as 128 x 64 bf16
// operand tiles in shared memory, produced by CUDA-core warps straight into the 128-byte-swizzled K-major layout the
// tensor core reads, while TMA streams the matching W tiles:
//   warps 0-7   generators : thread = (row, k-block parity).  The tile's X^0 rows (bf16 pairs) and X^{k-1} rows sit in shared
//                            memory ([m / 2][row], [h][row]: conflict-free, double-buffered, the next tile's values are
//                            loaded into registers while this one is generated).  Per 16-byte chunk: 4 x (LDS + one packed
//                            bf16 multiply), one 16-byte store at chunk ^ (row & 7); fence.proxy.async + one arrive per warp.
//   warps 8-11  epilogue   : thread = row.  bias + ReLU + bf16 rounding, hidden columns to global, direct columns dotted with
//                            pool_w in fp32, one atomicAdd per row into out_acc[b].
//   warp 12     TMA        : W tile [128 x 64] per k-block (rows >= O and columns >= H*M are zero-filled)
//   warp 13     MMA issuer : 4 x tcgen05.mma (M 128, N 128, K 16) per k-block into one of two TMEM accumulators
//   warp 14     TMEM alloc
// Rounding points are those of the unfused path (z rounded once to bf16: the packed bf16 multiply rounds the exact
// product; Y rounded to bf16 before pooling), so both give the same values up to fp32 accumulation order.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace oov {
namespace tc {

constexpr int C_BM = 128, C_BN = 128, C_BK = 64;
constexpr int C_ASTAGES = 4, C_BSTAGES = 5;                  // separate rings: generated z tiles / W tiles (TMA latency needs the depth)
constexpr int C_A_BYTES = C_BM * C_BK * 2, C_B_BYTES = C_BN * C_BK * 2;
constexpr int C_RING_BYTES = C_ASTAGES * C_A_BYTES + C_BSTAGES * C_B_BYTES;
constexpr int C_MMAX = 64, C_HMAX = 64;                       // fields / hidden channels of the previous layer
constexpr int C_GEN_WARPS = 8, C_EPI_WARP0 = 8, C_W_TMA = 12, C_W_MMA = 13, C_W_ALLOC = 14;
constexpr int C_THREADS = 15 * 32;
constexpr int C_X0_BYTES = 2 * (C_MMAX / 2) * C_BM * 4;       // [2][M / 2][128] bf16 pairs
constexpr int C_XROWS = C_HMAX + 32;                          // h of a padding channel reaches H + 62 / M <= H + 31 (rows >= H are zeros)
constexpr int C_XI_BYTES = 2 * C_XROWS * C_BM * 2;            // [2][C_XROWS][128] bf16
constexpr int C_SMEM = 1024 + C_RING_BYTES + C_X0_BYTES + C_XI_BYTES + 2 * C_BN * 4 + 256;
static_assert(C_SMEM <= 232448, "tc_cin shared memory");

struct CinParams {
    const __nv_bfloat16* x0; int64_t x0_sd; int M;                     // row r = (b, d) of X^0 at x0 + r * x0_sd, M channels (d-major)
    const __nv_bfloat16* xi; int64_t xi_sd; int H;                     // row r of X^{k-1} at xi + r * xi_sd, H channels
    int64_t B; int D; int64_t R;                                       // R = B * D rows
    int mp_shift;                                                      // z channel = (h << mp_shift) + m, 2^mp_shift >= max(M, 8)
    int k_blocks; int O;                                               // ceil((H << mp_shift) / 64); output channels (<= 128)
    const float* bias;                                                 // [O]
    __nv_bfloat16* hid_out; int64_t ld_h; int n_hidden;                // Y[:, :n_hidden] (even; 0: none)
    int pool_lo, pool_n; const float* pool_w; float* out_acc;          // direct-connect columns and their cin_linear weights
    int debug;                                                         // profiling only (OOV_CIN_DEBUG): 1 = generators only signal, 2 = W tile loaded once per stage slot
};

__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void gen_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(C_THREADS, 1)
tc_cin_layer_kernel(const __grid_constant__ CUtensorMap tmW, const CinParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                                                                   // C_ASTAGES x 16 KB
    unsigned char* sB = smem + C_ASTAGES * C_A_BYTES;                                           // C_BSTAGES x 16 KB
    uint32_t* x0s = reinterpret_cast<uint32_t*>(smem + C_RING_BYTES);                           // [2][32][128]
    uint16_t* xis = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(x0s) + C_X0_BYTES);   // [2][96][128]
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(xis) + C_XI_BYTES);      // [128]
    float* poolw_s = bias_s + C_BN;                                                            // [128] by output column
    uint64_t* a_full = reinterpret_cast<uint64_t*>(poolw_s + C_BN);
    uint64_t* a_empty = a_full + C_ASTAGES;
    uint64_t* b_full = a_empty + C_ASTAGES;
    uint64_t* b_empty = b_full + C_BSTAGES;
    uint64_t* tmem_full = b_empty + C_BSTAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.R + C_BM - 1) / C_BM;
    const int KB = p.k_blocks;

    if (warp == C_W_TMA && lane == 0) tma_prefetch_desc(&tmW);
    if (warp == C_W_MMA && lane == 0) {
        for (int s = 0; s < C_ASTAGES; ++s) { mbar_init(&a_full[s], C_GEN_WARPS / 2); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < C_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == C_W_ALLOC) tmem_alloc(tmem_slot, 2 * C_BN);
    for (int i = threadIdx.x; i < (C_X0_BYTES + C_XI_BYTES) / 4; i += C_THREADS) x0s[i] = 0u;   // fields >= M / channels >= H stay zero
    for (int i = threadIdx.x; i < C_BN; i += C_THREADS) {
        bias_s[i] = i < p.O ? __ldg(p.bias + i) : 0.f;
        const int c = i - p.pool_lo;
        poolw_s[i] = (c >= 0 && c < p.pool_n) ? __ldg(p.pool_w + c) : 0.f;          // zero weight outside the direct part
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < C_GEN_WARPS) {
        // ===================== generators =====================
        const int row = threadIdx.x & (C_BM - 1), half = warp >> 2;     // this thread: tile row, k-blocks kb = half (mod 2)
        const int M = p.M, H = p.H, M2 = M >> 1;
        const uint32_t mp_shift = (uint32_t)p.mp_shift, mp_mask = (1u << mp_shift) - 1u;
        constexpr int NX = C_MMAX / 4, NH = C_HMAX / 2;                 // values of a tile this thread fetches: pairs m2 = half + 2 i, h = half + 2 j
        uint32_t px[NX];
        // both operands are read with unit channel stride (rows of [R, ld] matrices, ld even): bf16 pairs as 32-bit loads, and
        // nothing touches the loaded registers before store_tile — the loads stay in flight for a whole tile
        uint32_t pw[NH / 2];
        auto load_tile = [&](int64_t tile) {
            const int64_t r = tile * C_BM + row;
            const bool ok = r < p.R;
            const uint32_t* q0 = reinterpret_cast<const uint32_t*>(p.x0 + (ok ? r : 0) * p.x0_sd);
            const uint32_t* qi = reinterpret_cast<const uint32_t*>(p.xi + (ok ? r : 0) * p.xi_sd);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const int m2 = half + 2 * i;
                px[i] = (ok && m2 < M2) ? __ldg(q0 + m2) : 0u;
            }
#pragma unroll
            for (int j = 0; j < NH / 2; ++j) {
                const int h2 = half + 2 * j;                            // channels 2 h2, 2 h2 + 1
                pw[j] = (ok && 2 * h2 < H) ? __ldg(qi + h2) : 0u;
            }
        };
        auto store_tile = [&](int buf) {
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const int m2 = half + 2 * i;
                if (m2 < M2) x0s[(buf * (C_MMAX / 2) + m2) * C_BM + row] = px[i];
            }
#pragma unroll
            for (int j = 0; j < NH / 2; ++j) {
                const int h2 = half + 2 * j;
                const uint32_t v = pw[j];
                xis[(buf * C_XROWS + 2 * h2) * C_BM + row] = (uint16_t)(v & 0xffffu);
                xis[(buf * C_XROWS + 2 * h2 + 1) * C_BM + row] = (2 * h2 + 1 < H) ? (uint16_t)(v >> 16) : (uint16_t)0;   // rows >= H are zeros
            }
        };
        int64_t t = blockIdx.x;
        if (t < n_tiles) load_tile(t);
        uint32_t s_cnt = 0;                                             // k-blocks of earlier tiles
        for (int it = 0; t < n_tiles; t += gridDim.x, ++it, s_cnt += (uint32_t)KB) {
            const int buf = it & 1;
            store_tile(buf);
            gen_bar();                                                  // the tile's rows are in shared memory (and buffer buf ^ 1 is free)
            if (t + gridDim.x < n_tiles) load_tile(t + gridDim.x);      // in flight while this tile is generated
            // 32-bit shared addresses of this thread's column of the two row buffers: x0 pair of fields (m, m + 1) at
            // x0a + (m << 8), X^{k-1} channel h at xia + (h << 8)
            const uint32_t x0a = smem_u32(x0s + buf * (C_MMAX / 2) * C_BM + row);
            const uint32_t xia = smem_u32(xis + buf * C_XROWS * C_BM + row);
            const uint32_t sw = (uint32_t)(row & 7);
            for (int kb = half; kb < KB; kb += 2) {
                const uint32_t s = s_cnt + (uint32_t)kb;
                const int stage = (int)(s % C_ASTAGES);
                const uint32_t phase = (s / C_ASTAGES) & 1u;
                mbar_wait(&a_empty[stage], phase ^ 1u);                 // the MMAs that read this stage have retired
                const uint32_t arow = smem_u32(sA + stage * C_A_BYTES + row * 128);
                // z channel c = h * Mp + m with Mp = the field count rounded up to a power of two (>= 8): a 16-byte chunk (8
                // channels) has ONE h, and h / m are a shift and a mask — per chunk one LDS.U16 + broadcast for X^{k-1}[h], per
                // pair one LDS (immediate offset) + one packed bf16 multiply.  (With c = h*M + m every pair needed its own
                // multiply-shift division and two loads: 8 instructions per pair, the generators were issue-bound at 30 % tensor
                // pipe.)  Fields m >= M are zero rows of the buffer and zero columns of W; so are channels h >= H.
                const uint32_t c0 = (uint32_t)(kb * C_BK);
                uint32_t av[2], xv[2][4];
                auto loads = [&](int j, uint32_t& a, uint32_t (&x)[4]) {
                    const uint32_t c = c0 + (uint32_t)(8 * j);
                    a = lds_u16(xia + ((c >> mp_shift) << 8));
                    const uint32_t xa = x0a + ((c & mp_mask) << 8);
#pragma unroll
                    for (int q = 0; q < 4; ++q) x[q] = lds_u32(xa + (uint32_t)(q * 512));
                };
                if (!(p.debug & 1)) loads(0, av[0], xv[0]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (p.debug & 1) break;
                    if (j + 1 < 8) loads(j + 1, av[(j + 1) & 1], xv[(j + 1) & 1]);
                    const uint32_t a2 = av[j & 1] * 0x10001u;
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[q] = bf16x2_mul(a2, xv[j & 1][q]);
                    sts_v4(arow + (((uint32_t)j ^ sw) << 4), w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();                               // generic-proxy stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
            }
        }
    } else if (warp == C_W_TMA) {
        // ===================== TMA: W tiles =====================
        const bool issue = elect_one();
        uint32_t s = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++s) {
                const int stage = (int)(s % C_BSTAGES);
                mbar_wait_relaxed(&b_empty[stage], ((s / C_BSTAGES) & 1u) ^ 1u);
                if (issue) {
                    if ((p.debug & 2) && s >= C_BSTAGES) mbar_arrive(&b_full[stage]);       // timing experiment: no load
                    else {
                        mbar_arrive_expect_tx(&b_full[stage], C_B_BYTES);
                        tma_load_2d(sB + stage * C_B_BYTES, &tmW, &b_full[stage], kb * C_BK, 0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == C_W_MMA) {
        // ===================== MMA issuer =====================
        const bool issue = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16_f32(C_BM, C_BN);
        uint32_t s = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C_BN);
            for (int kb = 0; kb < KB; ++kb, ++s) {
                const int sa_i = (int)(s % C_ASTAGES), sb_i = (int)(s % C_BSTAGES);
                mbar_wait(&b_full[sb_i], (s / C_BSTAGES) & 1u);         // W bytes landed
                mbar_wait(&a_full[sa_i], (s / C_ASTAGES) & 1u);         // all four generator warps of this k-block arrived
                tc_fence_after();
                const uint64_t adesc = make_sw128_desc(smem_u32(sA + sa_i * C_A_BYTES)), bdesc = make_sw128_desc(smem_u32(sB + sb_i * C_B_BYTES));
                if (issue) {
#pragma unroll
                    for (int k = 0; k < C_BK / 16; ++k)
                        if (!(p.debug & 8)) tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    tc_commit(&a_empty[sa_i]);
                    tc_commit(&b_empty[sb_i]);
                    if (kb == KB - 1) tc_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= C_EPI_WARP0 && warp < C_EPI_WARP0 + 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const int n_chunks = (p.O + 31) >> 5;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait_relaxed(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int64_t r = t * C_BM + q * 32 + lane;
            const bool ok = r < p.R;
            float dot = 0.f;
            __nv_bfloat16* hrow = p.hid_out ? p.hid_out + (ok ? r : 0) * p.ld_h : nullptr;
            for (int c4 = 0; c4 < n_chunks; ++c4) {
                if (p.debug & 4) break;                                 // timing experiment: no epilogue work
                uint32_t v[32];
                tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C_BN + c4 * 32), v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int col = c4 * 32 + j;
                    const float y0 = relu_nan(__uint_as_float(v[j]) + bias_s[col]);
                    const float y1 = relu_nan(__uint_as_float(v[j + 1]) + bias_s[col + 1]);
                    __nv_bfloat162 hp = __floats2bfloat162_rn(y0, y1);
                    if (ok && hrow != nullptr && col < p.n_hidden) *reinterpret_cast<__nv_bfloat162*>(hrow + col) = hp;
                    dot = fmaf(__bfloat162float(hp.x), poolw_s[col], dot);
                    dot = fmaf(__bfloat162float(hp.y), poolw_s[col + 1], dot);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (ok && p.pool_n > 0) atomicAdd(p.out_acc + r / p.D, dot);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == C_W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * C_BN);
    }
}

bool cin_tc_supported(int H, int M, int O, int n_hidden, int64_t ld_h) {
    return M >= 2 && M <= C_MMAX && (M & 1) == 0 && H >= 1 && H <= C_HMAX && O >= 1 && O <= C_BN && (n_hidden & 1) == 0 &&
           n_hidden <= O && (ld_h & 1) == 0;
}

int cin_tc_run(const CinParams& p, const void* W, int64_t ldw, cudaStream_t st) {
    CUtensorMap tmW;
    int rc = make_tmap_bf16_2d(&tmW, W, (uint64_t)ldw, (uint64_t)p.O, (uint64_t)ldw * 2, C_BN);   // columns [H*M, ldw) are zero padding
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(tc_cin_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM);
    OOV_REQUIRE(e == cudaSuccess, OOV_ERR_CUDA, "cudaFuncSetAttribute(tc_cin_layer_kernel): %s", cudaGetErrorString(e));
    const int64_t n_tiles = cdiv(p.R, C_BM);
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    tc_cin_layer_kernel<<<grid, C_THREADS, C_SMEM, st>>>(tmW, p);
    OOV_LAUNCH_CHECK("tc_cin_layer_kernel");
    return OOV_OK;
}

}  // namespace tc
}  // namespace oov

using namespace oov;

extern "C" {

int oov_cin_layer_supported(int32_t H, int32_t M, int32_t O, int32_t n_hidden, int64_t ld_h) {
    return tc::cin_tc_supported(H, M, O, n_hidden, ld_h) ? 1 : 0;
}

int oov_cin_layer(const void* xi, int64_t ld_xi, int32_t H, const void* x0, int64_t ld_x0, int32_t M, int32_t Mp,
                  int64_t B, int32_t D, const void* W, int64_t ldw, const float* bias, int32_t O,
                  void* hid_out, int64_t ld_h, int32_t n_hidden,
                  int32_t pool_lo, int32_t pool_n, const float* pool_w, float* out_acc, void* stream) {
    OOV_REQUIRE(B >= 0 && D > 0 && H > 0 && M > 0 && O > 0, OOV_ERR_ARG, "oov_cin_layer: bad shape");
    OOV_REQUIRE(tc::cin_tc_supported(H, M, O, n_hidden, ld_h), OOV_ERR_ARG,
                "oov_cin_layer: needs even M <= 64, H <= 64, O <= 128, even n_hidden / ld_h (H=%d M=%d O=%d); use oov_cin_outer + oov_tc_linear", H, M, O);
    OOV_REQUIRE(Mp >= 8 && Mp >= M && Mp <= 64 && (Mp & (Mp - 1)) == 0, OOV_ERR_ARG, "oov_cin_layer: Mp=%d must be a power of two in [max(M, 8), 64]", Mp);
    OOV_REQUIRE(ldw % 8 == 0 && ldw >= (int64_t)H * Mp, OOV_ERR_ALIGN, "oov_cin_layer: ldw must be a multiple of 8 >= H*Mp");
    OOV_REQUIRE(pool_lo >= 0 && pool_n >= 0 && pool_lo + pool_n <= O && n_hidden >= 0 && n_hidden <= O, OOV_ERR_ARG, "oov_cin_layer: bad column ranges");
    if (B == 0) return OOV_OK;
    OOV_REQUIRE(xi && x0 && W && bias && (n_hidden == 0 || hid_out) && (pool_n == 0 || (pool_w && out_acc)), OOV_ERR_ARG, "oov_cin_layer: NULL pointer");
    OOV_REQUIRE(n_hidden == 0 || (aligned(hid_out, 4) && ld_h >= n_hidden), OOV_ERR_ALIGN, "oov_cin_layer: hid_out must be 4-byte aligned with ld_h >= n_hidden");
    tc::CinParams p{};
    OOV_REQUIRE(aligned(xi, 4) && aligned(x0, 4) && ld_xi % 2 == 0 && ld_x0 % 2 == 0 && ld_xi >= H && ld_x0 >= M, OOV_ERR_ALIGN,
                "oov_cin_layer: xi / x0 must be 4-byte aligned [B*D, ld] matrices with even ld >= H / M");
    p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.x0_sd = ld_x0; p.M = M;
    p.xi = reinterpret_cast<const __nv_bfloat16*>(xi); p.xi_sd = ld_xi; p.H = H;
    p.B = B; p.D = D; p.R = B * (int64_t)D;
    p.mp_shift = 0;
    while ((1 << p.mp_shift) < Mp) ++p.mp_shift;
    p.k_blocks = (int)cdiv((int64_t)H * Mp, tc::C_BK); p.O = O; p.bias = bias;
    p.hid_out = n_hidden ? reinterpret_cast<__nv_bfloat16*>(hid_out) : nullptr; p.ld_h = ld_h; p.n_hidden = n_hidden;
    p.pool_lo = pool_lo; p.pool_n = pool_n; p.pool_w = pool_w; p.out_acc = out_acc;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("OOV_CIN_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    return tc::cin_tc_run(p, W, ldw, (cudaStream_t)stream);
}

}  // extern "C"
