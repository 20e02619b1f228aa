// DHE (deep hash embedding) — SipHash-2-4 multi-hash encoder + CUDA-core fp32 MLP path.
//
// Restates inductive/dh_embedder.py:140-170 (128 keyed SipHash-2-4 of the 8-byte LE id, % 2^24,
// held as exact fp32) and :70-89 (Linear-GELU x3, Linear-Sigmoid) of the reference.  The
// reference calls the csiphash C wheel once per (id, key) from a Python loop; here one thread
// computes one (id, key) hash with the per-key initial state staged in shared memory.
// The tensor-core MLP lives in tc_dhe.cu; this file holds the hash kernel and the fp32 path.
#include "common.cuh"

namespace oov {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

#define OOV_SIPROUND(v0, v1, v2, v3) \
    do {                             \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                      \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                      \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)

// SipHash-2-4 of one 8-byte message m given the key-derived initial state.
__device__ __forceinline__ uint64_t siphash24_8(uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3, uint64_t m) {
    v3 ^= m;
    OOV_SIPROUND(v0, v1, v2, v3);
    OOV_SIPROUND(v0, v1, v2, v3);
    v0 ^= m;
    const uint64_t b = 8ull << 56;           // length block: len = 8, no tail bytes
    v3 ^= b;
    OOV_SIPROUND(v0, v1, v2, v3);
    OOV_SIPROUND(v0, v1, v2, v3);
    v0 ^= b;
    v2 ^= 0xff;
    OOV_SIPROUND(v0, v1, v2, v3);
    OOV_SIPROUND(v0, v1, v2, v3);
    OOV_SIPROUND(v0, v1, v2, v3);
    OOV_SIPROUND(v0, v1, v2, v3);
    return v0 ^ v1 ^ v2 ^ v3;
}

__device__ __forceinline__ uint64_t load_le64(const uint8_t* p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

constexpr int HASH_THREADS = 256;

// One thread per (id, key); j (key) is the fastest index so the [n, H] output is coalesced.
__global__ void __launch_bounds__(HASH_THREADS)
dhe_hash_kernel(const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n,
                const uint8_t* __restrict__ keys, int H, uint64_t mod, uint32_t* __restrict__ out) {
    extern __shared__ uint64_t kst[];        // SoA: v0[H] | v1[H] | v2[H] | v3[H]
    for (int j = threadIdx.x; j < H; j += HASH_THREADS) {
        const uint64_t k0 = load_le64(keys + 16 * j), k1 = load_le64(keys + 16 * j + 8);
        kst[j] = k0 ^ 0x736f6d6570736575ull;
        kst[H + j] = k1 ^ 0x646f72616e646f6dull;
        kst[2 * H + j] = k0 ^ 0x6c7967656e657261ull;
        kst[3 * H + j] = k1 ^ 0x7465646279746573ull;
    }
    __syncthreads();
    const bool pow2 = (mod & (mod - 1)) == 0;
    const int64_t total = n * (int64_t)H;
    for (int64_t t = (int64_t)blockIdx.x * HASH_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * HASH_THREADS) {
        const int64_t i = t / H;
        const int j = (int)(t - i * H);
        const uint64_t m = (uint64_t)ids[i * ids_stride];   // LE bytes of the int64 id == its value
        const uint64_t h = siphash24_8(kst[j], kst[H + j], kst[2 * H + j], kst[3 * H + j], m);
        out[t] = (uint32_t)(pow2 ? (h & (mod - 1)) : (h % mod));
    }
}

// ------------------------------------------------------------------------------------
// fp32 linear + activation: C[n, N] = act(A[n, K] . W[N, K]^T + b)
// 64x64 CTA tile, 16-deep K slices, 256 threads, 4x4 micro-tile.
// ------------------------------------------------------------------------------------
enum { ACT_GELU = 0, ACT_SIGMOID = 1, ACT_NONE = 2 };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

template <bool A_IS_U32, int ACT>
__global__ void __launch_bounds__(256)
linear_act_simt(const void* __restrict__ A_, int64_t n, int K, const float* __restrict__ W, const float* __restrict__ bias,
                int N, float* __restrict__ C, int64_t ldc,
                // final-layer assemble (ACT_SIGMOID only): out rows + ids; C may be NULL then
                const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n_old,
                const void* __restrict__ iv_table, int iv_dtype, void* __restrict__ out, int out_dtype, int64_t out_stride) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += 16) {
        // 64 rows x 16 k = 1024 elements per operand, 4 per thread; k fastest in global
        for (int e = tid; e < 64 * 16; e += 256) {
            const int r = e >> 4, kk = e & 15;
            const int64_t gr = m0 + r;
            float av = 0.f;
            if (gr < n && k0 + kk < K) {
                if (A_IS_U32) av = (float)reinterpret_cast<const uint32_t*>(A_)[gr * K + k0 + kk];   // < 2^24: exact
                else av = reinterpret_cast<const float*>(A_)[gr * K + k0 + kk];
            }
            As[kk][r] = av;
            const int gn = n0 + r;
            Ws[kk][r] = (gn < N && k0 + kk < K) ? __ldg(W + (size_t)gn * K + k0 + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = m0 + ty * 4 + i;
        if (gr >= n) continue;
        int64_t id = 0;
        bool is_iv = false;
        if (ACT == ACT_SIGMOID && ids != nullptr) {
            id = ids[gr * ids_stride];
            is_iv = id < n_old;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + __ldg(bias + gn);
            v = (ACT == ACT_GELU) ? gelu_erf(v) : (ACT == ACT_SIGMOID ? sigmoidf(v) : v);
            if (ACT == ACT_SIGMOID && out != nullptr) {
                if (is_iv) {
                    if (iv_table == nullptr || id < 0) continue;
                    v = load_elem(iv_table, iv_dtype, id * (int64_t)N + gn);
                }
                store_elem(out, out_dtype, gr * out_stride + gn, v);
            } else {
                C[gr * ldc + gn] = v;
            }
        }
    }
}

int check_rows_public(const oov_rows* r, const char* who);
namespace tc {
size_t dhe_tc_workspace(int64_t n, const oov_dhe_net* net);
bool dhe_tc_supported(const oov_dhe_net* net, uint64_t mod);
int dhe_tc_run(const uint8_t* keys, uint64_t mod, const oov_dhe_net* net, const uint32_t* hashes_u32,
               const int64_t* ids, int64_t ids_stride, int64_t n, int64_t n_old, const void* iv_table, int iv_dtype,
               void* out, int out_dtype, int64_t out_stride, void* workspace, size_t workspace_bytes, cudaStream_t st,
               const __nv_bfloat16* planes_in = nullptr, const float* feat = nullptr, int64_t n_feat_rows = 0,
               int64_t prime_pad = 0);
int64_t dhe_planes_ld(int H);
int dhe_tc_hash_planes(const int64_t* ids, int64_t ids_stride, int64_t n, const uint8_t* keys, int H, uint64_t mod,
                       __nv_bfloat16* planes, cudaStream_t st);
}  // namespace tc
// OOV_PATH_AUTO: bf16 outputs take the tensor-core path (bf16 operands, fp32 accumulate, rtol 1e-3 contract),
// fp32 outputs take the CUDA-core fp32 path (rtol 1e-5 contract).
static bool use_tc(int path, int out_dtype, const oov_dhe_net* net, uint64_t mod) {
    if (path == OOV_PATH_SIMT_FP32) return false;
    if (!tc::dhe_tc_supported(net, mod)) return false;
    return path == OOV_PATH_TCGEN05 || out_dtype == OOV_BF16;
}
constexpr int64_t DHE_CHUNK = 1 << 16;     // rows per pass of the fp32 path (bounds the workspace)

static int check_net(const oov_dhe_net* net, const char* who, bool feat_ok = false) {
    OOV_REQUIRE(net != nullptr, OOV_ERR_ARG, "%s: net is NULL", who);
    OOV_REQUIRE(net->H >= 0 && net->F >= 0 && net->H + net->F > 0 && net->hidden > 0 && net->D > 0, OOV_ERR_ARG, "%s: bad net dims", who);
    OOV_REQUIRE(feat_ok || (net->F == 0 && net->H > 0), OOV_ERR_ARG, "%s: a net with feature inputs (F=%d) goes through oov_fdhe_embed", who, net->F);
    for (int l = 0; l < 4; ++l) OOV_REQUIRE(net->w[l] && net->b[l], OOV_ERR_ARG, "%s: NULL weight/bias %d", who, l);
    return OOV_OK;
}

static size_t dhe_simt_workspace(int64_t n, const oov_dhe_net* net) {
    const int64_t c = n < DHE_CHUNK ? n : DHE_CHUNK;
    return align_up((size_t)c * net->H * 4, 256) + 2 * align_up((size_t)c * net->hidden * 4, 256);
}

static int launch_hash(const int64_t* ids, int64_t ids_stride, int64_t n, const uint8_t* keys, int H, uint64_t mod,
                       uint32_t* out, cudaStream_t st) {
    if (n == 0) return OOV_OK;
    int64_t blocks = cdiv(n * (int64_t)H, HASH_THREADS);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    dhe_hash_kernel<<<(unsigned)blocks, HASH_THREADS, (size_t)H * 32, st>>>(ids, ids_stride, n, keys, H, mod, out);
    OOV_LAUNCH_CHECK("dhe_hash_kernel");
    return OOV_OK;
}

// hashes: [n, H] u32 already computed; rows (optional) gives the assemble contract for the last layer
// x_f32 (optional, [n, H + F] fp32) replaces the hashes as the first layer's input (fdhe / dnn)
static int mlp_simt(const uint32_t* hashes, int64_t n, const oov_dhe_net* net, const oov_rows* rows,
                    void* out, int out_dtype, int64_t out_stride, float* act0, float* act1, cudaStream_t st,
                    const float* x_f32 = nullptr) {
    const dim3 blk(256);
    const unsigned gy = (unsigned)cdiv(n, 64);
    const int hid = net->hidden;
    if (x_f32 != nullptr)
        linear_act_simt<false, ACT_GELU><<<dim3((unsigned)cdiv(hid, 64), gy), blk, 0, st>>>(
            x_f32, n, net->H + net->F, net->w[0], net->b[0], hid, act0, hid, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    else
    linear_act_simt<true, ACT_GELU><<<dim3((unsigned)cdiv(hid, 64), gy), blk, 0, st>>>(
        hashes, n, net->H, net->w[0], net->b[0], hid, act0, hid, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    OOV_LAUNCH_CHECK("linear_act_simt L1");
    linear_act_simt<false, ACT_GELU><<<dim3((unsigned)cdiv(hid, 64), gy), blk, 0, st>>>(
        act0, n, hid, net->w[1], net->b[1], hid, act1, hid, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    OOV_LAUNCH_CHECK("linear_act_simt L2");
    linear_act_simt<false, ACT_GELU><<<dim3((unsigned)cdiv(hid, 64), gy), blk, 0, st>>>(
        act1, n, hid, net->w[2], net->b[2], hid, act0, hid, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    OOV_LAUNCH_CHECK("linear_act_simt L3");
    linear_act_simt<false, ACT_SIGMOID><<<dim3((unsigned)cdiv(net->D, 64), gy), blk, 0, st>>>(
        act0, n, hid, net->w[3], net->b[3], net->D, nullptr, 0,
        rows ? rows->ids : nullptr, rows ? rows->ids_stride : 1, rows ? rows->n_old : 0,
        rows ? rows->iv_table : nullptr, rows ? rows->iv_dtype : 0, out, out_dtype, out_stride);
    OOV_LAUNCH_CHECK("linear_act_simt L4");
    return OOV_OK;
}

// fdhe / dnn first-layer input of the fp32 path: x[i] = [float(hash_0..H-1) | feat[id']]  (feat_dh_embedder.py:188-196)
__global__ void fdhe_input_kernel(const uint32_t* __restrict__ hashes, int H, const float* __restrict__ feat, int64_t n_feat_rows,
                                  int F, const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n, int64_t prime_pad,
                                  float* __restrict__ x) {
    const int K = H + F;
    const int64_t total = n * (int64_t)K;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / K;
        const int j = (int)(t - i * K);
        float v = 0.f;
        if (j < H) {
            v = (float)hashes[i * H + j];                       // < 2^24: exact
        } else {
            int64_t id = ids[i * ids_stride];
            if (prime_pad > 0 && id >= prime_pad) id -= prime_pad;
            if (id >= 0 && id < n_feat_rows) v = __ldg(feat + id * (int64_t)F + (j - H));
        }
        x[t] = v;
    }
}

// training: y = act(z) (dy == nullptr) or dz = dy * act'(z); rows whose id < n_old (in-vocab rows of an assembled batch,
// which take the table row instead of the net's output) get dz = 0.  act: 0 none, 1 GELU(erf), 2 sigmoid.
__global__ void act_kernel(const float* __restrict__ z, const float* __restrict__ dy, int act, int64_t rows, int N,
                           const int64_t* __restrict__ ids, int64_t ids_stride, int64_t n_old, float* __restrict__ out) {
    const int64_t total = rows * (int64_t)N;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const float x = z[t];
        float v;
        if (dy == nullptr) {
            v = act == 1 ? gelu_erf(x) : (act == 2 ? sigmoidf(x) : x);
        } else {
            float d = 1.f;
            if (act == 1) d = 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
            else if (act == 2) { const float s = sigmoidf(x); d = s * (1.f - s); }
            v = dy[t] * d;
            if (ids != nullptr && ids[(t / N) * ids_stride] < n_old) v = 0.f;
        }
        out[t] = v;
    }
}

static size_t fdhe_simt_workspace(int64_t n, const oov_dhe_net* net) {
    const int64_t c = n < DHE_CHUNK ? n : DHE_CHUNK;
    return align_up((size_t)c * (net->H > 0 ? net->H : 1) * 4, 256) + align_up((size_t)c * (net->H + net->F) * 4, 256) +
           2 * align_up((size_t)c * net->hidden * 4, 256);
}

}  // namespace oov

using namespace oov;

extern "C" {

int oov_dhe_hash(const int64_t* ids, int64_t ids_stride, int64_t n, const uint8_t* keys, int32_t H, uint64_t mod,
                 uint32_t* hashes, void* stream) {
    OOV_REQUIRE(n >= 0 && H > 0 && H <= 4096 && ids_stride >= 1, OOV_ERR_ARG, "oov_dhe_hash: bad shape n=%lld H=%d", (long long)n, H);
    OOV_REQUIRE(keys && (n == 0 || (ids && hashes)), OOV_ERR_ARG, "oov_dhe_hash: NULL pointer");
    OOV_REQUIRE(mod >= 1 && mod <= (1ull << 32), OOV_ERR_ARG, "oov_dhe_hash: mod must be in [1, 2^32]");
    return launch_hash(ids, ids_stride, n, keys, H, mod, hashes, (cudaStream_t)stream);
}

size_t oov_dhe_workspace(int64_t n, const oov_dhe_net* net, int32_t path) {
    if (!net || n <= 0) return 0;
    const size_t a = dhe_simt_workspace(n, net);
    if (path == OOV_PATH_SIMT_FP32 || !tc::dhe_tc_supported(net, 1ull << 24)) return a;
    const size_t b = tc::dhe_tc_workspace(n, net);      // AUTO may pick either path: size for the larger
    return a > b ? a : b;
}

int oov_dhe_mlp(const uint32_t* hashes, int64_t n, const oov_dhe_net* net, void* out, int32_t out_dtype,
                int64_t out_stride, void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    int rc = check_net(net, "oov_dhe_mlp");
    if (rc) return rc;
    OOV_REQUIRE(n >= 0 && dtype_ok(out_dtype) && out_stride >= net->D, OOV_ERR_ARG, "oov_dhe_mlp: bad n/out_dtype/out_stride");
    OOV_REQUIRE(path >= OOV_PATH_AUTO && path <= OOV_PATH_TCGEN05, OOV_ERR_ARG, "oov_dhe_mlp: unsupported path %d", path);
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(hashes && out, OOV_ERR_ARG, "oov_dhe_mlp: NULL pointer");
    OOV_REQUIRE(path != OOV_PATH_TCGEN05 || tc::dhe_tc_supported(net, 1ull << 24), OOV_ERR_ARG,
                "oov_dhe_mlp: net shape not supported by the tcgen05 path");
    if (use_tc(path, out_dtype, net, 1ull << 24))
        return tc::dhe_tc_run(nullptr, 1ull << 24, net, hashes, nullptr, 1, n, 0, nullptr, 0, out, out_dtype, out_stride,
                              workspace, workspace_bytes, (cudaStream_t)stream);
    const int64_t c = n < DHE_CHUNK ? n : DHE_CHUNK;
    const size_t hsz = align_up((size_t)c * net->H * 4, 256), asz = align_up((size_t)c * net->hidden * 4, 256);
    OOV_REQUIRE(workspace && workspace_bytes >= hsz + 2 * asz, OOV_ERR_WORKSPACE, "oov_dhe_mlp: workspace %zu < %zu",
                workspace_bytes, hsz + 2 * asz);
    float* act0 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + hsz);
    float* act1 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + hsz + asz);
    const size_t osz = dtype_size(out_dtype);
    for (int64_t r0 = 0; r0 < n; r0 += c) {
        const int64_t cn = n - r0 < c ? n - r0 : c;
        rc = mlp_simt(hashes + r0 * net->H, cn, net, nullptr, reinterpret_cast<char*>(out) + (size_t)r0 * out_stride * osz,
                      out_dtype, out_stride, act0, act1, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return OOV_OK;
}

int oov_dhe_embed(const uint8_t* keys, uint64_t mod, const oov_dhe_net* net, const oov_rows* rows,
                  void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    int rc = check_net(net, "oov_dhe_embed");
    if (rc) return rc;
    rc = check_rows_public(rows, "oov_dhe_embed");
    if (rc) return rc;
    OOV_REQUIRE(keys, OOV_ERR_ARG, "oov_dhe_embed: keys is NULL");
    OOV_REQUIRE(rows->D == net->D, OOV_ERR_ARG, "oov_dhe_embed: rows->D=%d != net->D=%d", rows->D, net->D);
    OOV_REQUIRE(mod >= 1 && mod <= (1ull << 32), OOV_ERR_ARG, "oov_dhe_embed: mod must be in [1, 2^32]");
    OOV_REQUIRE(path >= OOV_PATH_AUTO && path <= OOV_PATH_TCGEN05, OOV_ERR_ARG, "oov_dhe_embed: unsupported path %d", path);
    const int64_t n = rows->n;
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(path != OOV_PATH_TCGEN05 || tc::dhe_tc_supported(net, mod), OOV_ERR_ARG,
                "oov_dhe_embed: net shape / modulus not supported by the tcgen05 path");
    if (use_tc(path, rows->out_dtype, net, mod))
        return tc::dhe_tc_run(keys, mod, net, nullptr, rows->ids, rows->ids_stride, n, rows->n_old, rows->iv_table,
                              rows->iv_dtype, rows->out, rows->out_dtype, rows->out_stride, workspace, workspace_bytes,
                              (cudaStream_t)stream);
    const int64_t c = n < DHE_CHUNK ? n : DHE_CHUNK;
    const size_t hsz = align_up((size_t)c * net->H * 4, 256), asz = align_up((size_t)c * net->hidden * 4, 256);
    OOV_REQUIRE(workspace && workspace_bytes >= hsz + 2 * asz, OOV_ERR_WORKSPACE, "oov_dhe_embed: workspace %zu < %zu",
                workspace_bytes, hsz + 2 * asz);
    uint32_t* hashes = reinterpret_cast<uint32_t*>(workspace);
    float* act0 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + hsz);
    float* act1 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + hsz + asz);
    const size_t osz = dtype_size(rows->out_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t r0 = 0; r0 < n; r0 += c) {
        const int64_t cn = n - r0 < c ? n - r0 : c;
        oov_rows sub = *rows;
        sub.ids = rows->ids + r0 * rows->ids_stride;
        sub.n = cn;
        sub.out = reinterpret_cast<char*>(rows->out) + (size_t)r0 * rows->out_stride * osz;
        // dh_embedder.py:219-245: DHE hashes the raw id (prime pad NOT removed)
        rc = launch_hash(sub.ids, sub.ids_stride, cn, keys, net->H, mod, hashes, st);
        if (rc) return rc;
        rc = mlp_simt(hashes, cn, net, &sub, sub.out, sub.out_dtype, sub.out_stride, act0, act1, st);
        if (rc) return rc;
    }
    return OOV_OK;
}

int64_t oov_dhe_planes_ld(int32_t H) { return tc::dhe_planes_ld(H); }

int oov_dhe_hash_planes(const int64_t* ids, int64_t ids_stride, int64_t n, const uint8_t* keys, int32_t H, uint64_t mod,
                        void* planes, void* stream) {
    OOV_REQUIRE(keys && (n == 0 || (ids && planes)), OOV_ERR_ARG, "oov_dhe_hash_planes: NULL pointer");
    OOV_REQUIRE(H > 0 && n >= 0 && ids_stride >= 1, OOV_ERR_ARG, "oov_dhe_hash_planes: bad shape H=%d n=%lld", H, (long long)n);
    OOV_REQUIRE(mod >= 1 && mod <= (1ull << 24), OOV_ERR_ARG, "oov_dhe_hash_planes: mod must be in [1, 2^24] (three byte planes)");
    OOV_REQUIRE(aligned(planes, 16), OOV_ERR_ALIGN, "oov_dhe_hash_planes: planes must be 16-byte aligned");
    return tc::dhe_tc_hash_planes(ids, ids_stride, n, keys, H, mod, reinterpret_cast<__nv_bfloat16*>(planes), (cudaStream_t)stream);
}

int oov_dhe_embed_planes(const void* planes, const oov_dhe_net* net, const oov_rows* rows, void* workspace,
                         size_t workspace_bytes, void* stream) {
    int rc = check_net(net, "oov_dhe_embed_planes");
    if (rc) return rc;
    rc = check_rows_public(rows, "oov_dhe_embed_planes");
    if (rc) return rc;
    OOV_REQUIRE(planes && aligned(planes, 16), OOV_ERR_ARG, "oov_dhe_embed_planes: planes is NULL or not 16-byte aligned");
    OOV_REQUIRE(rows->D == net->D, OOV_ERR_ARG, "oov_dhe_embed_planes: rows->D=%d != net->D=%d", rows->D, net->D);
    OOV_REQUIRE(tc::dhe_tc_supported(net, 1ull << 24), OOV_ERR_ARG, "oov_dhe_embed_planes: net shape not supported by the tcgen05 path");
    if (rows->n == 0) return OOV_OK;
    return tc::dhe_tc_run(nullptr, 1ull << 24, net, nullptr, rows->ids, rows->ids_stride, rows->n, rows->n_old, rows->iv_table,
                          rows->iv_dtype, rows->out, rows->out_dtype, rows->out_stride, workspace, workspace_bytes,
                          (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(planes));
}

int oov_linear_f32(const float* A, const float* W, int64_t M, int32_t N, int32_t K, const float* bias, int32_t act, float* out,
                   void* stream) {
    OOV_REQUIRE(M >= 0 && N > 0 && K > 0 && act >= 0 && act <= 2, OOV_ERR_ARG, "oov_linear_f32: bad argument");
    if (M == 0) return OOV_OK;
    OOV_REQUIRE(A && W && bias && out, OOV_ERR_ARG, "oov_linear_f32: NULL pointer");
    const dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64)), blk(256);
    OOV_REQUIRE(grid.y <= 65535u, OOV_ERR_ARG, "oov_linear_f32: M=%lld too large (chunk the rows)", (long long)M);
    cudaStream_t st = (cudaStream_t)stream;
    if (act == 1)
        linear_act_simt<false, ACT_GELU><<<grid, blk, 0, st>>>(A, M, K, W, bias, N, out, N, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    else if (act == 2)
        linear_act_simt<false, ACT_SIGMOID><<<grid, blk, 0, st>>>(A, M, K, W, bias, N, out, N, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    else
        linear_act_simt<false, ACT_NONE><<<grid, blk, 0, st>>>(A, M, K, W, bias, N, out, N, nullptr, 1, 0, nullptr, 0, nullptr, 0, 0);
    OOV_LAUNCH_CHECK("linear_act_simt");
    return OOV_OK;
}

int oov_act(const float* z, const float* dy, int32_t act, int64_t rows, int32_t N, const int64_t* ids, int64_t ids_stride,
            int64_t n_old, float* out, void* stream) {
    OOV_REQUIRE(rows >= 0 && N > 0 && act >= 0 && act <= 2 && ids_stride >= 1, OOV_ERR_ARG, "oov_act: bad argument");
    if (rows == 0) return OOV_OK;
    OOV_REQUIRE(z && out, OOV_ERR_ARG, "oov_act: NULL pointer");
    int64_t blocks = cdiv(rows * (int64_t)N, 256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, dy, act, rows, N, ids, ids_stride, n_old, out);
    OOV_LAUNCH_CHECK("act_kernel");
    return OOV_OK;
}

size_t oov_fdhe_input_workspace(int64_t n, int32_t H) { return n > 0 ? align_up((size_t)n * (H > 0 ? H : 1) * 4, 256) : 0; }

int oov_fdhe_input(const uint8_t* keys, uint64_t mod, int32_t H, const float* feat, int64_t n_feat_rows, int32_t F,
                   const int64_t* ids, int64_t ids_stride, int64_t n, int64_t prime_pad, float* x, void* workspace,
                   size_t workspace_bytes, void* stream) {
    OOV_REQUIRE(n >= 0 && H >= 0 && F >= 0 && H + F > 0 && ids_stride >= 1, OOV_ERR_ARG, "oov_fdhe_input: bad shape");
    OOV_REQUIRE(mod >= 1 && mod <= (1ull << 24), OOV_ERR_ARG, "oov_fdhe_input: mod must be in [1, 2^24] (exact in fp32)");
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(ids && x && (H == 0 || keys) && (F == 0 || (feat && n_feat_rows > 0)), OOV_ERR_ARG, "oov_fdhe_input: NULL pointer");
    OOV_REQUIRE(workspace && workspace_bytes >= oov_fdhe_input_workspace(n, H), OOV_ERR_WORKSPACE, "oov_fdhe_input: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* hashes = reinterpret_cast<uint32_t*>(workspace);
    if (H > 0) {
        int rc = launch_hash(ids, ids_stride, n, keys, H, mod, hashes, st);
        if (rc) return rc;
    }
    int64_t blocks = cdiv(n * (int64_t)(H + F), 256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    fdhe_input_kernel<<<(unsigned)blocks, 256, 0, st>>>(hashes, H, feat, n_feat_rows, F, ids, ids_stride, n, prime_pad, x);
    OOV_LAUNCH_CHECK("fdhe_input_kernel");
    return OOV_OK;
}

size_t oov_fdhe_workspace(int64_t n, const oov_dhe_net* net, int32_t path) {
    if (!net || n <= 0) return 0;
    const size_t a = fdhe_simt_workspace(n, net);
    if (path == OOV_PATH_SIMT_FP32 || !tc::dhe_tc_supported(net, 1ull << 24)) return a;
    const size_t b = tc::dhe_tc_workspace(n, net);
    return a > b ? a : b;
}

int oov_fdhe_embed(const uint8_t* keys, uint64_t mod, const oov_dhe_net* net, const float* feat, int64_t n_feat_rows,
                   const oov_rows* rows, void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
    int rc = check_net(net, "oov_fdhe_embed", true);
    if (rc) return rc;
    rc = check_rows_public(rows, "oov_fdhe_embed");
    if (rc) return rc;
    OOV_REQUIRE(net->H == 0 || keys, OOV_ERR_ARG, "oov_fdhe_embed: keys is NULL but the net has %d hash inputs", net->H);
    OOV_REQUIRE(net->F == 0 || (feat && n_feat_rows > 0), OOV_ERR_ARG, "oov_fdhe_embed: feat is NULL / empty but the net has %d feature inputs", net->F);
    OOV_REQUIRE(rows->D == net->D, OOV_ERR_ARG, "oov_fdhe_embed: rows->D=%d != net->D=%d", rows->D, net->D);
    OOV_REQUIRE(mod >= 1 && mod <= (1ull << 32), OOV_ERR_ARG, "oov_fdhe_embed: mod must be in [1, 2^32]");
    OOV_REQUIRE(path >= OOV_PATH_AUTO && path <= OOV_PATH_TCGEN05, OOV_ERR_ARG, "oov_fdhe_embed: unsupported path %d", path);
    const int64_t n = rows->n;
    if (n == 0) return OOV_OK;
    OOV_REQUIRE(path != OOV_PATH_TCGEN05 || tc::dhe_tc_supported(net, mod), OOV_ERR_ARG,
                "oov_fdhe_embed: net shape / modulus not supported by the tcgen05 path");
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tc(path, rows->out_dtype, net, mod))
        return tc::dhe_tc_run(keys, mod, net, nullptr, rows->ids, rows->ids_stride, n, rows->n_old, rows->iv_table,
                              rows->iv_dtype, rows->out, rows->out_dtype, rows->out_stride, workspace, workspace_bytes, st,
                              nullptr, feat, n_feat_rows, rows->prime_pad);
    const int64_t c = n < DHE_CHUNK ? n : DHE_CHUNK;
    const int K = net->H + net->F;
    const size_t hsz = align_up((size_t)c * (net->H > 0 ? net->H : 1) * 4, 256), xsz = align_up((size_t)c * K * 4, 256),
                 asz = align_up((size_t)c * net->hidden * 4, 256);
    OOV_REQUIRE(workspace && workspace_bytes >= hsz + xsz + 2 * asz, OOV_ERR_WORKSPACE, "oov_fdhe_embed: workspace %zu < %zu",
                workspace_bytes, hsz + xsz + 2 * asz);
    char* ws = reinterpret_cast<char*>(workspace);
    uint32_t* hashes = reinterpret_cast<uint32_t*>(ws);
    float* x = reinterpret_cast<float*>(ws + hsz);
    float* act0 = reinterpret_cast<float*>(ws + hsz + xsz);
    float* act1 = reinterpret_cast<float*>(ws + hsz + xsz + asz);
    const size_t osz = dtype_size(rows->out_dtype);
    for (int64_t r0 = 0; r0 < n; r0 += c) {
        const int64_t cn = n - r0 < c ? n - r0 : c;
        oov_rows sub = *rows;
        sub.ids = rows->ids + r0 * rows->ids_stride;
        sub.n = cn;
        sub.out = reinterpret_cast<char*>(rows->out) + (size_t)r0 * rows->out_stride * osz;
        if (net->H > 0) {       // the hashes use the ORIGINAL id (feat_dh_embedder.py:199-205)
            rc = launch_hash(sub.ids, sub.ids_stride, cn, keys, net->H, mod, hashes, st);
            if (rc) return rc;
        }
        int64_t blocks = cdiv(cn * (int64_t)K, 256);
        const int64_t cap = (int64_t)num_sms() * 16;
        if (blocks > cap) blocks = cap;
        fdhe_input_kernel<<<(unsigned)blocks, 256, 0, st>>>(hashes, net->H, feat, n_feat_rows, net->F, sub.ids, sub.ids_stride, cn,
                                                           rows->prime_pad, x);
        OOV_LAUNCH_CHECK("fdhe_input_kernel");
        rc = mlp_simt(nullptr, cn, net, &sub, sub.out, sub.out_dtype, sub.out_stride, act0, act1, st, x);
        if (rc) return rc;
    }
    return OOV_OK;
}

}  // extern "C"
