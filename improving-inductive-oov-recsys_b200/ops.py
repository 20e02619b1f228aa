"""Tensor-level wrappers over the C-ABI (include/oov_b200.h).

torch is plumbing here: it owns device memory and the current stream; every op below
hands raw device pointers + sizes to liboov_b200.so and returns torch tensors.  All ops
are asynchronous on `torch.cuda.current_stream()` and never synchronise with the host.
They raise if the tensors are not on a CUDA device — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import INT64_MAX, OOV_BF16, OOV_F32, PATH_AUTO, PATH_SIMT_FP32, PATH_TCGEN05, OovDheNet, OovRows

TIE_EPS = 1e-6      # north_star: projection-sign ties with |x| < 1e-6 are counted and reported
MAX_HASH = 16777216  # reference dh_embedder.py:53


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return OOV_F32
    if t.dtype == torch.bfloat16:
        return OOV_BF16
    raise ValueError(f"unsupported dtype {t.dtype}: tables/outputs must be float32 or bfloat16")


def _torch_dtype(code_or_dtype) -> torch.dtype:
    if isinstance(code_or_dtype, torch.dtype):
        return code_or_dtype
    return torch.float32 if code_or_dtype == OOV_F32 else torch.bfloat16


def _cuda(t: Optional[torch.Tensor], name: str, dtype=None, allow_none=False) -> None:
    if t is None:
        if allow_none:
            return
        raise ValueError(f"{name} is None")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (got {t.device}); oov_b200 has no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype} (got {t.dtype})")


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache: dict = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand and reused (ordered by the stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _ids_1d(ids: torch.Tensor, name="ids") -> Tuple[torch.Tensor, int]:
    _cuda(ids, name, torch.int64)
    if ids.dim() != 1:
        raise ValueError(f"{name} must be 1-D (got shape {tuple(ids.shape)})")
    stride = ids.stride(0) if ids.numel() > 1 else 1
    if stride < 1:
        ids = ids.contiguous()
        stride = 1
    return ids, stride


def make_rows(ids: torch.Tensor, D: int, out: Optional[torch.Tensor] = None, out_dtype=torch.float32,
              n_old: int = 0, iv_table: Optional[torch.Tensor] = None, prime_pad: int = 0):
    """Build the `oov_rows` struct.  `out` may be a strided view [n, D] with unit inner stride
    (e.g. column f of a [B, fields, D] tensor); it is allocated when None."""
    ids, ids_stride = _ids_1d(ids)
    n = ids.shape[0]
    if out is None:
        out = torch.empty((n, D), dtype=_torch_dtype(out_dtype), device=ids.device)
    _cuda(out, "out")
    if out.dim() != 2 or out.shape[0] != n or out.shape[1] != D or (D > 1 and out.stride(1) != 1):
        raise ValueError(f"out must be [n={n}, D={D}] with unit inner stride (got {tuple(out.shape)}, strides {out.stride()})")
    out_stride = out.stride(0) if n > 1 else max(D, out.stride(0))
    if iv_table is not None:
        _cuda(iv_table, "iv_table")
        if iv_table.dim() != 2 or iv_table.shape[1] != D or not iv_table.is_contiguous():
            raise ValueError("iv_table must be contiguous [rows, D]")
        if iv_table.shape[0] < n_old:
            raise ValueError(f"iv_table has {iv_table.shape[0]} rows < n_old={n_old}")
    r = OovRows()
    r.ids, r.ids_stride, r.n, r.n_old, r.prime_pad = ids.data_ptr(), ids_stride, n, int(n_old), int(prime_pad)
    r.iv_table = 0 if iv_table is None else iv_table.data_ptr()
    r.iv_dtype = OOV_F32 if iv_table is None else _dt(iv_table)
    r.out_dtype, r.out, r.out_stride, r.D = _dt(out), out.data_ptr(), out_stride, D
    return r, out, (ids, iv_table)          # keep-alive tuple


def new_counter(device) -> torch.Tensor:
    """uint64 device counter (held in an int64 tensor) for the reported projection-sign ties."""
    return torch.zeros(1, dtype=torch.int64, device=device)


# ------------------------------------------------------------------------------------ LSH / SLSH
def _feat_planes(feat: torch.Tensor, planes: torch.Tensor):
    _cuda(feat, "feature_mat", torch.float32)
    _cuda(planes, "planes", torch.float32)
    if feat.dim() != 2 or planes.dim() != 2 or feat.shape[1] != planes.shape[1]:
        raise ValueError(f"feature_mat {tuple(feat.shape)} / planes {tuple(planes.shape)} mismatch")
    return feat.contiguous(), planes.contiguous()


def lsh_bits(feat, planes, ids, prime_pad: int = 0, tie_eps: float = TIE_EPS,
             tie_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packed multi-hot words int32 [n, ceil(B/32)] (bit b&31 of word b>>5 = plane b).
    reference: TorchLSHash.hash_points (torch_hash.py:55-60) on feature_mat[ids]."""
    feat, planes = _feat_planes(feat, planes)
    ids, stride = _ids_1d(ids)
    B, F = planes.shape
    bits = torch.empty((ids.shape[0], (B + 31) // 32), dtype=torch.int32, device=feat.device)
    _lib.check(_lib.load().oov_lsh_bits(_p(feat), feat.shape[0], F, _p(planes), B, _p(ids), stride, ids.shape[0],
                                        int(prime_pad), float(tie_eps), _p(bits), _p(tie_count), PATH_AUTO, _stream()))
    return bits


def lsh_embed(feat, planes, oov_weight, ids, out=None, out_dtype=torch.float32, n_old: int = 0, iv_table=None,
              prime_pad: int = 0, tie_eps: float = TIE_EPS, tie_count=None, return_bits: bool = False,
              path: int = PATH_AUTO, side_cast=None):
    """(H @ W) / H.sum(1) for OOV rows (lsh_embedder.py:141-179) + in-vocab gather (bpr.py:94-125).
    `side_cast=(src fp32, dst bf16)`: contiguous tensors of equal size (a multiple of 8 elements); dst = bf16(src) is
    done inside the same launch (the in-vocab half of a bf16 item table, see oov_lsh_embed_cast)."""
    feat, planes = _feat_planes(feat, planes)
    _cuda(oov_weight, "oov_weight")
    oov_weight = oov_weight.contiguous()
    B, F = planes.shape
    if oov_weight.dim() != 2 or oov_weight.shape[0] != B:
        raise ValueError(f"oov_weight must be [B={B}, D] (got {tuple(oov_weight.shape)})")
    D = oov_weight.shape[1]
    rows, out, keep = make_rows(ids, D, out, out_dtype, n_old, iv_table, prime_pad)
    lib = _lib.load()
    bits = torch.empty((rows.n, (B + 31) // 32), dtype=torch.int32, device=feat.device) if return_bits else None
    ws_bytes = lib.oov_lsh_embed_workspace(rows.n, B, D, path)
    ws = _workspace(ws_bytes, feat.device) if ws_bytes else None
    csrc = cdst = None
    if side_cast is not None:
        csrc, cdst = side_cast
        _cuda(csrc, "side_cast src", torch.float32)
        _cuda(cdst, "side_cast dst", torch.bfloat16)
        if not (csrc.is_contiguous() and cdst.is_contiguous()) or csrc.numel() != cdst.numel() or csrc.numel() % 8:
            raise ValueError("side_cast needs contiguous fp32 / bf16 tensors of equal size, a multiple of 8 elements")
    _lib.check(lib.oov_lsh_embed_cast(_p(feat), feat.shape[0], F, _p(planes), B, _p(oov_weight), _dt(oov_weight),
                                      C.byref(rows), float(tie_eps), _p(bits), _p(tie_count), _p(ws),
                                      0 if ws is None else ws.numel(), path, _p(csrc), _p(cdst),
                                      0 if csrc is None else csrc.numel(), _stream()))
    return (out, bits) if return_bits else out


def slsh_embed(feat, planes, n_buckets: int, oov_weight, ids, out=None, out_dtype=torch.float32, n_old: int = 0,
               iv_table=None, prime_pad: int = 0, tie_eps: float = TIE_EPS, tie_count=None,
               return_buckets: bool = False):
    """bucket = (bits_req + popcount(bits)) % n_buckets; out = W[bucket]
    (single_lsh_embedder.py:82-109).  `oov_weight=None` computes bucket ids only."""
    feat, planes = _feat_planes(feat, planes)
    bits_req, F = planes.shape
    if bits_req == 0:
        # n_buckets == 1: ceil(log2(1)) = 0 planes, every id lands in bucket (0 + 0) % 1 = 0 like the reference
        # (single_lsh_embedder.py:77-87); an empty tensor has a NULL data pointer, the C-ABI wants a valid one
        planes = torch.zeros((1, F), dtype=torch.float32, device=feat.device)
    lib = _lib.load()
    buckets = None
    if oov_weight is None:
        ids, stride = _ids_1d(ids)
        rows = OovRows()
        rows.ids, rows.ids_stride, rows.n, rows.n_old, rows.prime_pad = ids.data_ptr(), stride, ids.shape[0], int(n_old), int(prime_pad)
        rows.D = 1
        out = None
        return_buckets = True
        wptr, wdt = None, OOV_F32
    else:
        _cuda(oov_weight, "oov_weight")
        oov_weight = oov_weight.contiguous()
        if oov_weight.dim() != 2 or oov_weight.shape[0] != n_buckets:
            raise ValueError(f"oov_weight must be [n_buckets={n_buckets}, D] (got {tuple(oov_weight.shape)})")
        rows, out, keep = make_rows(ids, oov_weight.shape[1], out, out_dtype, n_old, iv_table, prime_pad)
        wptr, wdt = oov_weight, _dt(oov_weight)
    if return_buckets:
        buckets = torch.empty((rows.n,), dtype=torch.int64, device=feat.device)
    _lib.check(lib.oov_slsh_embed(_p(feat), feat.shape[0], F, _p(planes), bits_req, int(n_buckets), _p(wptr), wdt,
                                  C.byref(rows), float(tie_eps), _p(buckets), _p(tie_count), _stream()))
    if oov_weight is None:
        return buckets
    return (out, buckets) if return_buckets else out


# ------------------------------------------------------------------------------------ DHE
class DheNet:
    """Device view of one 4-layer hash net (dh_embedder.py:70-89): fp32 nn.Linear weights [out, in]."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], n_feat: int = 0):
        """`n_feat`: the last n_feat inputs of layer 1 are plain features (fdhe: feat_dh_embedder.py:100-101; dnn:
        dnn_embedder.py:65-66 with no hash inputs at all), the first H = in - n_feat are hash values."""
        if len(weights) != 4 or len(biases) != 4:
            raise ValueError("DHE net has exactly 4 Linear layers")
        self.w = [w.detach().contiguous() for w in weights]
        self.b = [b.detach().contiguous() for b in biases]
        for t in self.w + self.b:
            _cuda(t, "dhe weight", torch.float32)
        self.F = int(n_feat)
        self.H, self.hidden, self.D = self.w[0].shape[1] - self.F, self.w[0].shape[0], self.w[3].shape[0]
        if self.H < 0:
            raise ValueError("n_feat exceeds the first layer's input width")
        if self.w[1].shape != (self.hidden, self.hidden) or self.w[2].shape != (self.hidden, self.hidden) \
                or self.w[3].shape[1] != self.hidden:
            raise ValueError("DHE net layer shapes are inconsistent")
        s = OovDheNet()
        for l in range(4):
            s.w[l] = self.w[l].data_ptr()
            s.b[l] = self.b[l].data_ptr()
        s.H, s.hidden, s.D, s.F = self.H, self.hidden, self.D, self.F
        self.struct = s

    @classmethod
    def from_sequential(cls, net: torch.nn.Sequential, n_feat: int = 0) -> "DheNet":
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        return cls([m.weight for m in lin], [m.bias for m in lin], n_feat)


def keys_tensor(keys: Sequence[bytes], device) -> torch.Tensor:
    """list of 16-byte keys (dh_embedder.py:95-120) -> uint8 [H, 16] on the device."""
    if any(len(k) != 16 for k in keys):
        raise ValueError("SipHash keys must be 16 bytes")
    return torch.tensor(list(b"".join(keys)), dtype=torch.uint8).view(len(keys), 16).to(device)


def dhe_hash(ids, keys: torch.Tensor, mod: int = MAX_HASH) -> torch.Tensor:
    """int32 [n, H]: LE_u64(SipHash-2-4(key_j, LE8(id_i))) % mod  (dh_embedder.py:140-170)."""
    ids, stride = _ids_1d(ids)
    _cuda(keys, "keys", torch.uint8)
    keys = keys.contiguous()
    out = torch.empty((ids.shape[0], keys.shape[0]), dtype=torch.int32, device=ids.device)
    _lib.check(_lib.load().oov_dhe_hash(_p(ids), stride, ids.shape[0], _p(keys), keys.shape[0], int(mod), _p(out), _stream()))
    return out


def dhe_mlp(hashes: torch.Tensor, net: DheNet, out=None, out_dtype=torch.float32, path: int = PATH_AUTO):
    _cuda(hashes, "hashes", torch.int32)
    hashes = hashes.contiguous()
    n = hashes.shape[0]
    if hashes.dim() != 2 or hashes.shape[1] != net.H:
        raise ValueError(f"hashes must be [n, H={net.H}]")
    if out is None:
        out = torch.empty((n, net.D), dtype=_torch_dtype(out_dtype), device=hashes.device)
    lib = _lib.load()
    ws = _workspace(lib.oov_dhe_workspace(n, C.byref(net.struct), path), hashes.device)
    _lib.check(lib.oov_dhe_mlp(_p(hashes), n, C.byref(net.struct), _p(out), _dt(out), out.stride(0) if n > 1 else net.D,
                               _p(ws), ws.numel(), path, _stream()))
    return out


def dhe_embed(ids, keys: torch.Tensor, net: DheNet, out=None, out_dtype=torch.float32, n_old: int = 0, iv_table=None,
              mod: int = MAX_HASH, path: int = PATH_AUTO):
    """hash + MLP (+ in-vocab gather).  DHE does not de-pad ids (dh_embedder.py:219-245)."""
    _cuda(keys, "keys", torch.uint8)
    keys = keys.contiguous()
    if keys.shape[0] != net.H:
        raise ValueError(f"{keys.shape[0]} keys but the net expects H={net.H}")
    rows, out, keep = make_rows(ids, net.D, out, out_dtype, n_old, iv_table, 0)
    lib = _lib.load()
    ws = _workspace(lib.oov_dhe_workspace(rows.n, C.byref(net.struct), path), keys.device)
    _lib.check(lib.oov_dhe_embed(_p(keys), int(mod), C.byref(net.struct), C.byref(rows), _p(ws), ws.numel(), path, _stream()))
    return out


def dhe_hash_planes(ids, keys: torch.Tensor, mod: int = MAX_HASH) -> torch.Tensor:
    """bf16 [n, ld]: the three exact byte planes of every 24-bit hash (h = 65536 a + 256 b + c), the layout the
    tensor-core MLP consumes.  The hashes depend on (id, keys) only — the reference memoises them per id
    (dh_embedder.py:139 `@cache`) — so a caller that embeds the same ids again keeps this tensor."""
    ids, stride = _ids_1d(ids)
    _cuda(keys, "keys", torch.uint8)
    keys = keys.contiguous()
    lib = _lib.load()
    ld = int(lib.oov_dhe_planes_ld(keys.shape[0]))
    planes = torch.empty((ids.shape[0], ld), dtype=torch.bfloat16, device=ids.device)
    _lib.check(lib.oov_dhe_hash_planes(_p(ids), stride, ids.shape[0], _p(keys), keys.shape[0], int(mod), _p(planes), _stream()))
    return planes


def dhe_embed_planes(planes: torch.Tensor, ids, net: DheNet, out=None, out_dtype=torch.bfloat16, n_old: int = 0, iv_table=None):
    """MLP + assemble from memoised byte planes (`dhe_hash_planes` of the same ids)."""
    _cuda(planes, "planes", torch.bfloat16)
    rows, out, keep = make_rows(ids, net.D, out, out_dtype, n_old, iv_table, 0)
    lib = _lib.load()
    if planes.dim() != 2 or planes.shape[0] != rows.n or planes.shape[1] != int(lib.oov_dhe_planes_ld(net.H)) or not planes.is_contiguous():
        raise ValueError(f"planes must be contiguous [n={rows.n}, {int(lib.oov_dhe_planes_ld(net.H))}] (got {tuple(planes.shape)})")
    ws = _workspace(lib.oov_dhe_workspace(rows.n, C.byref(net.struct), PATH_TCGEN05), planes.device)
    _lib.check(lib.oov_dhe_embed_planes(_p(planes), C.byref(net.struct), C.byref(rows), _p(ws), ws.numel(), _stream()))
    return out


def fdhe_embed(ids, keys: Optional[torch.Tensor], net: DheNet, feat: Optional[torch.Tensor], out=None, out_dtype=torch.float32,
               n_old: int = 0, iv_table=None, prime_pad: int = 0, mod: int = MAX_HASH, path: int = PATH_AUTO):
    """`fdhe` (feat_dh_embedder.py:188-213) and, with net.H == 0, `dnn` (dnn_embedder.py:87-107): the 4-layer net on
    [hashes of the raw id | feature row of the de-padded id] (+ in-vocab gather).  `prime_pad` > 0 = training mode."""
    if net.H:
        _cuda(keys, "keys", torch.uint8)
        keys = keys.contiguous()
        if keys.shape[0] != net.H:
            raise ValueError(f"{keys.shape[0]} keys but the net expects H={net.H}")
    else:
        keys = None
    if net.F:
        _cuda(feat, "feat", torch.float32)
        feat = feat.contiguous()
        if feat.dim() != 2 or feat.shape[1] != net.F:
            raise ValueError(f"feat must be [rows, F={net.F}] (got {tuple(feat.shape)})")
    else:
        feat = None
    rows, out, keep = make_rows(ids, net.D, out, out_dtype, n_old, iv_table, prime_pad)
    lib = _lib.load()
    ws = _workspace(lib.oov_fdhe_workspace(rows.n, C.byref(net.struct), path), out.device)
    _lib.check(lib.oov_fdhe_embed(_p(keys), int(mod), C.byref(net.struct), _p(feat), feat.shape[0] if feat is not None else 0,
                                  C.byref(rows), _p(ws), ws.numel(), path, _stream()))
    return out


def tc_linear(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, act: str = "none",
              out_dtype=torch.float32, _debug: int = 0) -> torch.Tensor:
    """act(A @ W.T + bias) on the tensor cores: A [M, K], W [N, K] bf16, fp32 accumulate (tcgen05 + TMEM)."""
    _cuda(A, "A", torch.bfloat16)
    _cuda(W, "W", torch.bfloat16)
    _cuda(bias, "bias", torch.float32, allow_none=True)
    A, W = A.contiguous(), W.contiguous()
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise ValueError("A / W inner dimensions differ")
    out = torch.empty((M, N), dtype=_torch_dtype(out_dtype), device=A.device)
    code = {"none": 0, "gelu": 1, "sigmoid": 2, "relu": 3}[act] | (_debug << 8)
    _lib.check(_lib.load().oov_tc_linear(_p(A), K, _p(W), K, M, N, K, _p(bias), code, _p(out), _dt(out), N, _stream()))
    return out


# ------------------------------------------------------------------------------------ mean / zero / gathers
def col_mean(table: torch.Tensor) -> torch.Tensor:
    """fp32 [D] mean over ALL rows (mean_embedder.py:55-60)."""
    _cuda(table, "table")
    table = table.contiguous()
    rows_, D = table.shape
    lib = _lib.load()
    ws = _workspace(lib.oov_col_mean_workspace(rows_, D), table.device)
    out = torch.empty((D,), dtype=torch.float32, device=table.device)
    _lib.check(lib.oov_col_mean(_p(table), _dt(table), rows_, D, _p(out), _p(ws), ws.numel(), _stream()))
    return out


def const_embed(vec: Optional[torch.Tensor], ids, D: int, out=None, out_dtype=torch.float32, n_old: int = 0, iv_table=None):
    """OOV rows get `vec` (None = zeros): MeanEmbedder / ZeroEmbedder + in-vocab gather."""
    if vec is not None:
        _cuda(vec, "vec", torch.float32)
        vec = vec.contiguous()
        if vec.numel() != D:
            raise ValueError("vec must have D elements")
    rows, out, keep = make_rows(ids, D, out, out_dtype, n_old, iv_table, 0)
    _lib.check(_lib.load().oov_const_embed(_p(vec), C.byref(rows), _stream()))
    return out


def gather_rows(table: torch.Tensor, idx, idx_offset: int = 0, out=None, out_dtype=None):
    _cuda(table, "table")
    table = table.contiguous()
    idx, stride = _ids_1d(idx, "idx")
    n, D = idx.shape[0], table.shape[1]
    if out is None:
        out = torch.empty((n, D), dtype=table.dtype if out_dtype is None else _torch_dtype(out_dtype), device=table.device)
    else:
        _cuda(out, "out")
        _dt(out)
        if out.device != table.device or out.dim() != 2 or out.shape[0] != n or out.shape[1] != D or (D > 1 and out.stride(1) != 1):
            raise ValueError(f"out must be [n={n}, D={D}] on {table.device} with unit inner stride "
                             f"(got {tuple(out.shape)}, strides {out.stride()}, {out.device})")
    _lib.check(_lib.load().oov_gather_rows(_p(table), _dt(table), table.shape[0], D, _p(idx), stride, n, int(idx_offset),
                                           _p(out), _dt(out), out.stride(0) if n > 1 else D, _stream()))
    return out


def map_ids(ids, n_old: int, n_buckets: int, fn: str) -> torch.Tensor:
    code = {"mod": 0, "fast": 1, "3round": 2, "64bit": 3}.get(fn)
    if code is None:
        raise ValueError(f"Unknown hash function {fn}")
    _cuda(ids, "ids", torch.int64)
    ids = ids.contiguous()
    out = torch.empty_like(ids)
    _lib.check(_lib.load().oov_map_ids(_p(ids), ids.numel(), int(n_old), int(n_buckets), code, _p(out), _stream()))
    return out


# ------------------------------------------------------------------------------------ scoring + top-k
def _hist(hist, Q, device):
    if hist is None:
        return None, None
    rowptr, cols = hist
    _cuda(rowptr, "hist_rowptr", torch.int32)
    _cuda(cols, "hist_cols", torch.int32)
    if rowptr.numel() != Q + 1:
        raise ValueError(f"hist_rowptr must have Q+1={Q + 1} entries")
    if cols.numel() == 0:                      # keep a valid pointer for the NULL-consistency check
        cols = torch.zeros(1, dtype=torch.int32, device=device)
    return rowptr.contiguous(), cols.contiguous()


def _pad16(users: torch.Tensor, items: torch.Tensor):
    """The scoring kernels read rows with 128-bit loads: zero-pad D to a multiple of 16 (zeros add
    nothing to a dot product).  A copy — only taken for embedding sizes that are not 16-aligned."""
    D = users.shape[1]
    if D % 16 == 0:
        return users, items
    pad = 16 - D % 16
    return torch.nn.functional.pad(users, (0, pad)), torch.nn.functional.pad(items, (0, pad))


def fullsort_topk(users: torch.Tensor, items: torch.Tensor, k: int, item_id_offset: int = 0, mask_pad: bool = True,
                  seg: Tuple[int, int] = (0, INT64_MAX), hist=None, path: int = PATH_AUTO):
    """Fused score + mask + top-k: returns (scores fp32 [Q,k], global ids int64 [Q,k]) ordered by
    (score desc, id asc).  reference: bpr.py:151-156 + evaluator.py:91-94 + collector.py:153-159."""
    _cuda(users, "users")
    _cuda(items, "items")
    if users.dtype != items.dtype:
        raise ValueError("users and items must share a dtype")
    users, items = users.contiguous(), items.contiguous()
    Q, D = users.shape
    N = items.shape[0]
    if N and items.shape[1] != D:
        raise ValueError("users / items embedding size mismatch")
    users, items = _pad16(users, items)
    D = users.shape[1]
    rowptr, cols = _hist(hist, Q, users.device)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=users.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=users.device)
    lib = _lib.load()
    ws = _workspace(lib.oov_fullsort_topk_workspace(Q, max(N, 1), D, k, path), users.device)
    _lib.check(lib.oov_fullsort_topk(_p(users), _p(items), _dt(users), Q, N, D, k, int(item_id_offset), int(bool(mask_pad)),
                                     int(seg[0]), int(seg[1]), _p(rowptr), _p(cols), _p(out_s), _p(out_i), _p(ws),
                                     ws.numel(), path, _stream()))
    return out_s, out_i


def fullsort_topk_keys(users: torch.Tensor, items: torch.Tensor, k: int, row_map: Tuple[int, int, int], mask_pad: bool = True,
                       seg: Tuple[int, int] = (0, INT64_MAX), hist=None, path: int = PATH_AUTO) -> torch.Tensor:
    """`fullsort_topk` over a shard table laid out [ids lo0.. (n0 rows) | ids lo1..] (row_map = (n0, lo0, lo1)) that
    returns the lists as packed 8-byte candidates in GLOBAL ids: int64 [Q, k] holding uint64 keys
    (score bits << 32 | ~id; 0 = empty), the payload of the row-shard all-gather.  `seg` and `hist` are in local rows."""
    _cuda(users, "users")
    _cuda(items, "items")
    if users.dtype != items.dtype:
        raise ValueError("users and items must share a dtype")
    users, items = users.contiguous(), items.contiguous()
    Q, D = users.shape
    N = items.shape[0]
    if N and items.shape[1] != D:
        raise ValueError("users / items embedding size mismatch")
    users, items = _pad16(users, items)
    D = users.shape[1]
    rowptr, cols = _hist(hist, Q, users.device)
    keys = torch.empty((Q, k), dtype=torch.int64, device=users.device)
    lib = _lib.load()
    ws = _workspace(lib.oov_fullsort_topk_workspace(Q, max(N, 1), D, k, path), users.device)
    n0, lo0, lo1 = (int(v) for v in row_map)
    _lib.check(lib.oov_fullsort_topk_keys(_p(users), _p(items), _dt(users), Q, N, D, k, int(bool(mask_pad)), int(seg[0]), int(seg[1]),
                                          _p(rowptr), _p(cols), n0, lo0, lo1, _p(keys), _p(ws), ws.numel(), path, _stream()))
    return keys


def topk_merge_keys(keys: torch.Tensor):
    """[G, Q, k] packed candidates (`fullsort_topk_keys` of G shards) -> global (scores fp32, ids int64) [Q, k]."""
    _cuda(keys, "keys", torch.int64)
    keys = keys.contiguous()
    G, Q, k = keys.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=keys.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=keys.device)
    _lib.check(_lib.load().oov_topk_merge_keys(_p(keys), G, Q, k, _p(out_s), _p(out_i), _stream()))
    return out_s, out_i


def fullsort_scores(users, items, item_id_offset: int = 0, mask_pad: bool = False, seg=(0, INT64_MAX), hist=None):
    """Dense fp32 [Q, N] scores (what the reference materialises); masks optional."""
    _cuda(users, "users")
    _cuda(items, "items")
    users, items = users.contiguous(), items.contiguous()
    Q, D = users.shape
    N = items.shape[0]
    users, items = _pad16(users, items)
    D = users.shape[1]
    rowptr, cols = _hist(hist, Q, users.device)
    scores = torch.empty((Q, N), dtype=torch.float32, device=users.device)
    _lib.check(_lib.load().oov_fullsort_scores(_p(users), _p(items), _dt(users), Q, N, D, int(item_id_offset),
                                               int(bool(mask_pad)), int(seg[0]), int(seg[1]), _p(rowptr), _p(cols),
                                               _p(scores), N, _stream()))
    return scores


def dense_topk(scores: torch.Tensor, k: int):
    """torch.topk(scores, k) replacement for a materialised fp32 [Q, N] matrix (collector.py:153-159)."""
    _cuda(scores, "scores", torch.float32)
    if scores.dim() != 2 or scores.stride(1) != 1:
        scores = scores.contiguous()
    Q, N = scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    _lib.check(_lib.load().oov_dense_topk(_p(scores), scores.stride(0) if Q > 1 else N, Q, N, k, _p(out_s), _p(out_i), _stream()))
    return out_s, out_i


def topk_merge(cand_scores: torch.Tensor, cand_idx: torch.Tensor):
    """[G, Q, k] shard candidates -> global (scores, ids) [Q, k]."""
    _cuda(cand_scores, "cand_scores", torch.float32)
    _cuda(cand_idx, "cand_idx", torch.int64)
    cand_scores, cand_idx = cand_scores.contiguous(), cand_idx.contiguous()
    G, Q, k = cand_scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=cand_scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=cand_scores.device)
    _lib.check(_lib.load().oov_topk_merge(_p(cand_scores), _p(cand_idx), G, Q, k, _p(out_s), _p(out_i), _stream()))
    return out_s, out_i


def topk_hits(topk_idx: torch.Tensor, pos_rowptr: torch.Tensor, pos_cols: torch.Tensor) -> torch.Tensor:
    """'rec.topk' matrix [Q, k+1] = [hits | pos_len] (collector.py:160-166)."""
    _cuda(topk_idx, "topk_idx", torch.int64)
    _cuda(pos_rowptr, "pos_rowptr", torch.int32)
    _cuda(pos_cols, "pos_cols", torch.int32)
    topk_idx = topk_idx.contiguous()
    Q, k = topk_idx.shape
    if pos_cols.numel() == 0:
        pos_cols = torch.zeros(1, dtype=torch.int32, device=topk_idx.device)
    out = torch.empty((Q, k + 1), dtype=torch.int32, device=topk_idx.device)
    _lib.check(_lib.load().oov_topk_hits(_p(topk_idx), Q, k, _p(pos_rowptr.contiguous()), _p(pos_cols.contiguous()), _p(out), _stream()))
    return out


def topk_hits_collectors(idx_all, idx_old, idx_new, user_ids, n_old_users: int, n_old_items: int, pos_rowptr, pos_cols,
                         reference_compat: bool = False) -> torch.Tensor:
    """int32 [7, Q, k + 2] = hits | pos_len | keep for the seven collectors of inductive/evaluator.py:29-56 (order of
    evaluator.COLLECTORS) from the all / old-items / new-items k-lists of one scoring pass.  No host sync."""
    for t, nm in ((idx_all, "idx_all"), (idx_old, "idx_old"), (idx_new, "idx_new"), (user_ids, "user_ids")):
        _cuda(t, nm, torch.int64)
    _cuda(pos_rowptr, "pos_rowptr", torch.int32)
    _cuda(pos_cols, "pos_cols", torch.int32)
    idx_all, idx_old, idx_new, user_ids = idx_all.contiguous(), idx_old.contiguous(), idx_new.contiguous(), user_ids.contiguous()
    Q, k = idx_all.shape
    if idx_old.shape != (Q, k) or idx_new.shape != (Q, k) or user_ids.shape != (Q,) or pos_rowptr.numel() != Q + 1:
        raise ValueError("topk_hits_collectors: the three lists must be [Q, k], user_ids [Q], pos_rowptr [Q + 1]")
    if pos_cols.numel() == 0:
        pos_cols = torch.zeros(1, dtype=torch.int32, device=idx_all.device)
    out = torch.empty((7, Q, k + 2), dtype=torch.int32, device=idx_all.device)
    _lib.check(_lib.load().oov_topk_hits_collectors(_p(idx_all), _p(idx_old), _p(idx_new), Q, k, _p(user_ids), int(n_old_users),
                                                    int(n_old_items), _p(pos_rowptr.contiguous()), _p(pos_cols.contiguous()),
                                                    int(bool(reference_compat)), _p(out), _stream()))
    return out


def pairs_to_csr(rows_idx: torch.Tensor, cols_idx: torch.Tensor, Q: int, col_ranges=None, zero_tail: bool = False):
    """(row, item) index pairs (general_dataloader.py:270-292 history_index / positive_u,i) -> CSR
    (int32 rowptr [Q + 1], int32 cols ascending per row).  Rows outside [0, Q) are padding and are dropped; cols has
    the length of the input and is only meaningful up to rowptr[Q].  One kernel, no host sync (graph-capturable), any Q
    and up to 2^31 - 1 pairs (ceil(Q / 512) CTAs, each streams the pair list).  CUDA tensors only.

    col_ranges = ((lo0, hi0), (lo1, hi1)): keep only items of these two id ranges and rewrite them as LOCAL rows of a
    shard table laid out [range 0 | range 1] (sharded.py)."""
    if rows_idx is None or rows_idx.numel() == 0:
        dev = rows_idx.device if rows_idx is not None else torch.device("cuda")
        if dev.type != "cuda":
            raise RuntimeError("pairs_to_csr: index tensors must live on a CUDA device; oov_b200 has no CPU fallback")
        return torch.zeros(Q + 1, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    n = rows_idx.numel()
    _cuda(rows_idx, "rows_idx", torch.int64)
    _cuda(cols_idx, "cols_idx", torch.int64)
    rows_idx, cols_idx = rows_idx.contiguous(), cols_idx.contiguous()
    rowptr = torch.empty(Q + 1, dtype=torch.int32, device=rows_idx.device)
    cols = (torch.zeros if zero_tail else torch.empty)(n, dtype=torch.int32, device=rows_idx.device)   # zero_tail: entries past rowptr[Q] are 0
    cr = None
    if col_ranges is not None:
        (a0, b0), (a1, b1) = col_ranges
        cr = (C.c_int64 * 4)(int(a0), int(b0), int(a1), int(b1))
    _lib.check(_lib.load().oov_pairs_to_csr(_p(rows_idx), _p(cols_idx), n, Q, cr, _p(rowptr), _p(cols), _stream()))
    return rowptr, cols


def pair_topk(user_e: torch.Tensor, item_e: torch.Tensor, rowptr: torch.Tensor, cols: torch.Tensor, k: int,
              seg: Optional[Tuple[int, int]] = None, keys: Optional[torch.Tensor] = None, normalize: bool = False):
    """Sampled-candidate evaluation (trainer.py:547-564 / inductive/evaluator.py:116-133 without the [users, N] matrix):
    row r's top-k (score desc, id asc) among its distinct candidates cols[rowptr[r]:rowptr[r+1]] with score
    user_e[r] . item_e[j] (fp32; `normalize`: both rows L2-normalised first, DirectAU.predict), optionally only ids in
    seg = (lo, hi).  Missing slots are (-inf, -1).
    Returns (scores [U, k] fp32, ids [U, k] int64, keys) — pass `keys` back in to select another segment of the same
    batch without recomputing the scores."""
    _cuda(user_e, "user_e")
    _cuda(item_e, "item_e")
    _cuda(rowptr, "rowptr", torch.int32)
    _cuda(cols, "cols", torch.int32)
    user_e, item_e, rowptr, cols = user_e.contiguous(), item_e.contiguous(), rowptr.contiguous(), cols.contiguous()
    U, D = user_e.shape
    n = cols.numel()
    if rowptr.numel() != U + 1 or item_e.shape != (n, D):
        raise ValueError("pair_topk: rowptr must be [U + 1] and item_e [len(cols), D]")
    compute = keys is None
    if compute:
        keys = torch.empty((max(n, 1),), dtype=torch.int64, device=user_e.device)
    elif keys.numel() < n or keys.dtype != torch.int64:
        raise ValueError("pair_topk: keys must be the int64 tensor a previous call returned")
    lo, hi = (0, _lib.INT64_MAX) if seg is None else (int(seg[0]), int(seg[1]))
    out_s = torch.empty((U, k), dtype=torch.float32, device=user_e.device)
    out_i = torch.empty((U, k), dtype=torch.int64, device=user_e.device)
    _lib.check(_lib.load().oov_pair_topk(_p(user_e), _dt(user_e), _p(item_e), _dt(item_e), D, _p(rowptr), _p(cols), U, n, int(bool(normalize)), int(k),
                                         lo, hi, _p(keys), int(compute), _p(out_s), _p(out_i), _stream()))
    return out_s, out_i, keys


# ------------------------------------------------------------------------------------ context models
def token_gather(tokens: torch.Tensor, offsets: torch.Tensor, table: torch.Tensor, n_users: int, n_items: int,
                 user_const=None, item_const=None, uid_idx: int = 0, iid_idx: int = 1, out=None, out_dtype=None):
    """[B, fields] ids -> [B, fields, D]; OOV user/item cells get the constants when given, else are
    left for a following *_embed call (abstract_recommender.py:794-842)."""
    _cuda(tokens, "token_fields", torch.int64)
    _cuda(offsets, "offsets", torch.int64)
    _cuda(table, "table")
    tokens, offsets, table = tokens.contiguous(), offsets.contiguous(), table.contiguous()
    Bn, fields = tokens.shape
    D = table.shape[1]
    if out is None:
        out = torch.empty((Bn, fields, D), dtype=table.dtype if out_dtype is None else _torch_dtype(out_dtype), device=table.device)
    for v, nm in ((user_const, "user_const"), (item_const, "item_const")):
        _cuda(v, nm, torch.float32, allow_none=True)
    _lib.check(_lib.load().oov_token_gather(_p(tokens), Bn, fields, _p(offsets), _p(table), _dt(table), table.shape[0], D,
                                            int(n_users), int(n_items), uid_idx, iid_idx, _p(user_const), _p(item_const),
                                            _p(out), _dt(out), _stream()))
    return out


def first_order_sum(tokens, offsets, table1, n_users: int, n_items: int, oov_user_val=None, oov_item_val=None,
                    uid_idx: int = 0, iid_idx: int = 1) -> torch.Tensor:
    """[B] = sum over fields of the D=1 table with OOV cells replaced (layers.py:1634-1693)."""
    _cuda(tokens, "token_fields", torch.int64)
    _cuda(table1, "table1", torch.float32)
    tokens, offsets = tokens.contiguous(), offsets.contiguous()
    t1 = table1.contiguous().view(-1)
    Bn, fields = tokens.shape
    out = torch.empty((Bn,), dtype=torch.float32, device=tokens.device)
    _lib.check(_lib.load().oov_first_order_sum(_p(tokens), Bn, fields, _p(offsets), _p(t1), t1.numel(), int(n_users),
                                               int(n_items), uid_idx, iid_idx, _p(oov_user_val), _p(oov_item_val),
                                               _p(out), _stream()))
    return out


def cross_update(x0: torch.Tensor, t: torch.Tensor, xl: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x0 * t + xl elementwise (bf16): the tail of a DCN-V2 cross layer, dcnv2.py:137-142."""
    for a, nm in ((x0, "x0"), (t, "t"), (xl, "xl")):
        _cuda(a, nm, torch.bfloat16)
    if x0.shape != t.shape or x0.shape != xl.shape:
        raise ValueError("cross_update: shapes differ")
    x0, t, xl = x0.contiguous(), t.contiguous(), xl.contiguous()
    if out is None:
        out = torch.empty_like(x0)
    _lib.check(_lib.load().oov_cross_update(_p(x0), _p(t), _p(xl), x0.numel(), _p(out), _stream()))
    return out


def cin_outer(xi: torch.Tensor, x0: torch.Tensor, D: int, first: bool) -> torch.Tensor:
    """z [B*D, align8(H*M)] bf16 with z[(b, d), h*M + m] = xi[b, h, d] * x0[b, m, d] (xdeepfm.py:160-165).
    x0: the gathered embeddings [B, M, D]; xi: x0 itself (`first`) or a CIN layer's hidden channels, a column slice
    [B*D, H] of the previous linear's output."""
    _cuda(x0, "x0", torch.bfloat16)
    _cuda(xi, "xi", torch.bfloat16)
    if x0.dim() != 3 or x0.shape[2] != D or not x0.is_contiguous():
        raise ValueError("x0 must be contiguous [B, M, D]")
    B, M = x0.shape[0], x0.shape[1]
    if first:
        H, vi = M, (M * D, 1, D)
    else:
        if xi.dim() != 2 or xi.shape[0] != B * D or xi.stride(1) != 1:
            raise ValueError("xi must be [B*D, H] with unit inner stride")
        H, vi = xi.shape[1], (D * xi.stride(0), xi.stride(0), 1)
    ldz = (H * M + 7) // 8 * 8
    z = torch.empty((B * D, ldz), dtype=torch.bfloat16, device=x0.device)
    _lib.check(_lib.load().oov_cin_outer(_p(xi), vi[0], vi[1], vi[2], H, _p(x0), M * D, 1, D, M, B, D, _p(z), ldz, _stream()))
    return z


def cin_layer_supported(H: int, M: int, O: int, n_hidden: int, ld_h: int) -> bool:
    return bool(_lib.load().oov_cin_layer_supported(int(H), int(M), int(O), int(n_hidden), int(ld_h)))


def cin_field_pitch(M: int) -> int:
    """Mp of oov_cin_layer: the field count rounded up to a power of two, at least 8."""
    p = 8
    while p < M:
        p *= 2
    return p


def cin_pack_weight(w: torch.Tensor, H: int, M: int, rows: int) -> torch.Tensor:
    """conv1d weight [O, H*M] fp32 -> bf16 [rows, align8(H*Mp)] in the channel layout of oov_cin_layer (column h*Mp + m)."""
    O, Mp = w.shape[0], cin_field_pitch(M)
    out = torch.zeros((rows, H, Mp), dtype=torch.float32, device=w.device)
    out[:O, :, :M] = w.float().reshape(O, H, M)
    out = out.reshape(rows, H * Mp)
    return torch.nn.functional.pad(out, (0, (-out.shape[1]) % 8)).to(torch.bfloat16).contiguous()


def cin_layer(xi: torch.Tensor, x0t: torch.Tensor, B: int, D: int, W: torch.Tensor, bias: torch.Tensor, n_hidden: int,
              pool_lo: int, pool_n: int, pool_w: torch.Tensor, out_acc: torch.Tensor) -> Optional[torch.Tensor]:
    """One CIN layer fused (oov_cin_layer): returns the hidden channels [B*D, ld_h] bf16 (first n_hidden columns valid;
    None when n_hidden == 0) and ADDS the pooled direct-connect part into out_acc [B] fp32.
    x0t [B*D, M]: the embeddings with rows (b, d) and the fields along the row; xi [B*D, H] (unit inner stride): the
    previous layer's hidden channels (x0t itself for the first layer)."""
    for t, nm in ((x0t, "x0t"), (xi, "xi"), (W, "W")):
        _cuda(t, nm, torch.bfloat16)
    for t, nm in ((bias, "bias"), (pool_w, "pool_w"), (out_acc, "out_acc")):
        _cuda(t, nm, torch.float32)
    if x0t.dim() != 2 or xi.dim() != 2 or x0t.shape[0] != B * D or xi.shape[0] != B * D or x0t.stride(1) != 1 or xi.stride(1) != 1:
        raise ValueError("x0t / xi must be [B*D, channels] with unit inner stride")
    M, H, O = x0t.shape[1], xi.shape[1], W.shape[0]
    Mp = cin_field_pitch(M)
    if not W.is_contiguous() or W.shape[1] < H * Mp or bias.numel() != O or pool_w.numel() != pool_n or not pool_w.is_contiguous() \
            or out_acc.numel() != B or not out_acc.is_contiguous():
        raise ValueError("cin_layer: W [O, >= H*Mp] contiguous (cin_pack_weight), bias [O], pool_w [pool_n], out_acc [B]")
    ld_h = (n_hidden + 7) // 8 * 8
    hid = torch.empty((B * D, ld_h), dtype=torch.bfloat16, device=x0t.device) if n_hidden else None
    _lib.check(_lib.load().oov_cin_layer(_p(xi), xi.stride(0), H, _p(x0t), x0t.stride(0), M, Mp, B, D, _p(W), W.shape[1], _p(bias.contiguous()), O,
                                         _p(hid), ld_h, n_hidden, pool_lo, pool_n, _p(pool_w), _p(out_acc), _stream()))
    return hid


def cin_pool_dot(y: torch.Tensor, col0: int, ncols: int, B: int, D: int, w: torch.Tensor, bias: float,
                 acc: Optional[torch.Tensor] = None, accumulate: bool = False):
    """acc[b] (+)= bias + sum_{d, c} y[(b, d), col0 + c] * w[c]  (xdeepfm.py:188-189 + :198); fp32 [B]."""
    _cuda(y, "y", torch.bfloat16)
    _cuda(w, "w", torch.float32)
    if y.dim() != 2 or y.shape[0] != B * D or y.stride(1) != 1 or w.numel() != ncols or not w.is_contiguous():
        raise ValueError("y must be [B*D, >= col0 + ncols] with unit inner stride, w contiguous [ncols]")
    if acc is None:
        if accumulate:
            raise ValueError("accumulate needs an existing acc")
        acc = torch.empty((B,), dtype=torch.float32, device=y.device)
    _cuda(acc, "acc", torch.float32)
    if acc.numel() != B or not acc.is_contiguous():
        raise ValueError("acc must be contiguous fp32 [B]")
    _lib.check(_lib.load().oov_cin_pool_dot(_p(y), y.stride(0), col0, ncols, B, D, _p(w), float(bias), int(accumulate), _p(acc), _stream()))
    return acc


def _grad2d(g: torch.Tensor, D: int) -> torch.Tensor:
    _cuda(g, "grad", torch.float32)
    if g.dim() != 2 or g.shape[1] != D:
        raise ValueError(f"grad must be fp32 [n, D={D}]")
    return g if g.stride(1) == 1 else g.contiguous()


def scatter_add_rows(g: torch.Tensor, idx: torch.Tensor, dtable: torch.Tensor, idx_offset: int = 0) -> torch.Tensor:
    """dtable[idx[i] + idx_offset] += g[i] for rows that land inside dtable (nn.Embedding backward; bpr.py:56-71)."""
    _cuda(dtable, "dtable", torch.float32)
    if dtable.dim() != 2 or not dtable.is_contiguous():
        raise ValueError("dtable must be contiguous fp32 [rows, D]")
    g = _grad2d(g, dtable.shape[1])
    idx, stride = _ids_1d(idx, "idx")
    if idx.shape[0] != g.shape[0]:
        raise ValueError("idx / grad row counts differ")
    _lib.check(_lib.load().oov_scatter_add_rows(_p(g), g.stride(0) if g.shape[0] > 1 else dtable.shape[1], _p(idx), stride, idx.shape[0],
                                                int(idx_offset), dtable.shape[0], dtable.shape[1], _p(dtable), _stream()))
    return dtable


def lsh_embed_backward(bits: torch.Tensor, g: torch.Tensor, ids: torch.Tensor, n_old: int, dW: torch.Tensor) -> torch.Tensor:
    """dW[b] += H_ib g_i / |H_i| over the OOV rows (ids >= n_old) of an `lsh_embed(..., return_bits=True)` call."""
    _cuda(bits, "bits", torch.int32)
    _cuda(dW, "dW", torch.float32)
    if dW.dim() != 2 or not dW.is_contiguous():
        raise ValueError("dW must be contiguous fp32 [B, D]")
    B, D = dW.shape
    g = _grad2d(g, D)
    ids, stride = _ids_1d(ids)
    n = ids.shape[0]
    if bits.shape != (n, (B + 31) // 32) or not bits.is_contiguous() or g.shape[0] != n:
        raise ValueError("bits must be contiguous [n, ceil(B / 32)] and grad [n, D]")
    lib = _lib.load()
    ws = _workspace(lib.oov_lsh_embed_backward_workspace(n), dW.device) if n else None
    _lib.check(lib.oov_lsh_embed_backward(_p(bits), B, _p(g), g.stride(0) if n > 1 else D, _p(ids), stride, n, int(n_old), D, _p(dW),
                                          _p(ws), 0 if ws is None else ws.numel(), _stream()))
    return dW


_ACT = {"none": 0, "gelu": 1, "sigmoid": 2}


def fdhe_input(ids, keys: Optional[torch.Tensor], feat: Optional[torch.Tensor], prime_pad: int = 0, mod: int = MAX_HASH) -> torch.Tensor:
    """fp32 [n, H + F]: hashes of the raw id (as floats) | feature row of the de-padded id — layer 1's input in training."""
    ids, stride = _ids_1d(ids)
    H = 0 if keys is None else keys.shape[0]
    F = 0 if feat is None else feat.shape[1]
    if keys is not None:
        _cuda(keys, "keys", torch.uint8)
        keys = keys.contiguous()
    if feat is not None:
        _cuda(feat, "feat", torch.float32)
        feat = feat.contiguous()
    n = ids.shape[0]
    x = torch.empty((n, H + F), dtype=torch.float32, device=ids.device)
    lib = _lib.load()
    ws = _workspace(lib.oov_fdhe_input_workspace(n, H), ids.device) if n else None
    _lib.check(lib.oov_fdhe_input(_p(keys), int(mod), H, _p(feat), 0 if feat is None else feat.shape[0], F, _p(ids), stride, n,
                                  int(prime_pad), _p(x), _p(ws), 0 if ws is None else ws.numel(), _stream()))
    return x


def linear_f32(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, act: str = "none") -> torch.Tensor:
    """fp32 act(A @ W.T + bias) on the CUDA cores (training path; A [M, K], W [N, K])."""
    _cuda(A, "A", torch.float32)
    _cuda(W, "W", torch.float32)
    A, W = A.contiguous(), W.contiguous()
    if A.dim() != 2 or W.dim() != 2 or A.shape[1] != W.shape[1]:
        raise ValueError("A [M, K] / W [N, K] inner dimensions differ")
    if bias is None:
        bias = torch.zeros((W.shape[0],), dtype=torch.float32, device=A.device)
    _cuda(bias, "bias", torch.float32)
    out = torch.empty((A.shape[0], W.shape[0]), dtype=torch.float32, device=A.device)
    _lib.check(_lib.load().oov_linear_f32(_p(A), _p(W), A.shape[0], W.shape[0], A.shape[1], _p(bias.contiguous()), _ACT[act], _p(out), _stream()))
    return out


def act_forward(z: torch.Tensor, act: str) -> torch.Tensor:
    _cuda(z, "z", torch.float32)
    z = z.contiguous()
    out = torch.empty_like(z)
    _lib.check(_lib.load().oov_act(_p(z), None, _ACT[act], z.shape[0], z.shape[1], None, 1, 0, _p(out), _stream()))
    return out


def act_backward(z: torch.Tensor, dy: torch.Tensor, act: str, ids: Optional[torch.Tensor] = None, n_old: int = 0) -> torch.Tensor:
    """dy * act'(z); with `ids`, rows whose id < n_old are zeroed."""
    _cuda(z, "z", torch.float32)
    _cuda(dy, "dy", torch.float32)
    z, dy = z.contiguous(), dy.contiguous()
    if z.shape != dy.shape:
        raise ValueError("z / dy shapes differ")
    stride = 1
    if ids is not None:
        ids, stride = _ids_1d(ids)
    out = torch.empty_like(z)
    _lib.check(_lib.load().oov_act(_p(z), _p(dy), _ACT[act], z.shape[0], z.shape[1], _p(ids), stride, int(n_old), _p(out), _stream()))
    return out


def launch_count() -> int:
    return int(_lib.load().oov_launch_count())
