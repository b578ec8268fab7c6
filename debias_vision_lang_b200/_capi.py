"""ctypes binding of libdvl_b200.so (include/dvl_b200.h).

The CUDA library is the product; there is no other implementation behind these calls.  If the
shared object is missing or no sm_100 device is present the calls raise -- loudly -- instead of
computing anything on the host.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DVL_LIB_PATH") or os.path.join(_HERE, "libdvl_b200.so")   # override: A/B builds only

DVL_OK = 0
DVL_F32, DVL_BF16, DVL_F16 = 0, 1, 2
PRECISION_FP32, PRECISION_BF16, PRECISION_FP32X3 = 0, 1, 2
FLAG_NORMALIZE_QUERIES, FLAG_NORMALIZE_GALLERY, FLAG_FULL_SORT, FLAG_STREAM_TOPK, FLAG_NO_PILOT = 1, 2, 4, 8, 16
FLAG_RADIX_SORT = 32
FLAG_REUSE_GALLERY = 64
MAX_CLASSES = 16
STAGES = ("prep", "gemm", "sort_hist", "sort_scan", "sort_scatter", "metrics", "select", "copy")
NUM_STAGES = len(STAGES)

EXPORTS = (
    "dvl_abi_version", "dvl_create", "dvl_destroy", "dvl_last_error", "dvl_workspace_bytes",
    "dvl_launch_count", "dvl_set_profiling", "dvl_get_profile", "dvl_encode_labels", "dvl_eval_ranking", "dvl_eval_ranking_host", "dvl_rank_and_score", "dvl_similarity",
    "dvl_shard_topk", "dvl_shard_topk_host", "dvl_shard_stats", "dvl_shard_reduce", "dvl_shard_trim_pack",
    "dvl_shard_finish", "dvl_shard_trim", "dvl_shard_pack", "dvl_merge_topk", "dvl_mean",
)

_STATUS = {-1: "DVL_ERR_INVALID", -2: "DVL_ERR_NO_DEVICE", -3: "DVL_ERR_CUDA", -4: "DVL_ERR_OOM",
           -5: "DVL_ERR_UNSUPPORTED", -6: "DVL_ERR_OVERFLOW"}


class DvlError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_STATUS.get(code, code)}: {message}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Loads the CUDA library (building is ``__graft_entry__.build()``'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    c_void_p, c_int, c_i32, c_i64, c_u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32
    L.dvl_abi_version.restype = c_int
    L.dvl_create.argtypes = [ctypes.POINTER(c_void_p), c_int]
    L.dvl_destroy.argtypes = [c_void_p]
    L.dvl_last_error.argtypes = [c_void_p]
    L.dvl_last_error.restype = ctypes.c_char_p
    L.dvl_workspace_bytes.argtypes = [c_void_p]
    L.dvl_workspace_bytes.restype = ctypes.c_size_t
    L.dvl_launch_count.argtypes = [c_void_p]
    L.dvl_launch_count.restype = ctypes.c_uint64
    L.dvl_set_profiling.argtypes = [c_void_p, c_int]
    L.dvl_get_profile.argtypes = [c_void_p, c_void_p, c_void_p]
    L.dvl_encode_labels.argtypes = [c_void_p, c_void_p, c_int, c_i64, c_void_p, c_void_p, c_void_p]
    L.dvl_eval_ranking.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_i64, c_i32, c_i32, c_i32,
                                   c_i64, c_int, c_u32, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p,
                                   c_void_p]
    L.dvl_eval_ranking_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_i64, c_i32, c_i32,
                                        c_i32, c_i64, c_int, c_u32, c_void_p, c_void_p, c_void_p]
    L.dvl_rank_and_score.argtypes = [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i32, c_i32, c_i64, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_i64, c_void_p]
    L.dvl_similarity.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_i64, c_i32, c_i32, c_int, c_u32,
                                 c_void_p, c_i64, c_void_p]
    L.dvl_shard_topk.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_i64, c_i64, c_i32, c_i32, c_i32,
                                 c_i64, c_int, c_u32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.dvl_shard_topk_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_i64, c_i64, c_i32, c_i32,
                                      c_i32, c_i64, c_int, c_u32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]
    L.dvl_shard_stats.argtypes = [c_void_p, c_void_p, c_i32, c_void_p]
    L.dvl_shard_reduce.argtypes = [c_void_p, c_void_p, c_i32, c_i32, c_void_p]
    L.dvl_shard_trim_pack.argtypes = [c_void_p, c_i32, c_i64, c_void_p, c_void_p, c_void_p]
    L.dvl_shard_finish.argtypes = [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_void_p, c_void_p, c_void_p]
    L.dvl_shard_trim.argtypes = [c_void_p, c_void_p, c_i32, c_void_p, c_void_p]
    L.dvl_shard_pack.argtypes = [c_void_p, c_void_p, c_i32, c_i64, c_void_p, c_void_p]
    L.dvl_merge_topk.argtypes = [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_i32, c_i64, c_void_p, c_i32, c_i64,
                                 c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]
    L.dvl_mean.argtypes = [c_void_p, c_void_p, c_i32, c_void_p, c_void_p]
    for name in EXPORTS:
        if name not in ("dvl_last_error", "dvl_workspace_bytes", "dvl_launch_count"):
            getattr(L, name).restype = c_int
    _lib = L
    return L


def _dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return DVL_F32
    if t == torch.bfloat16:
        return DVL_BF16
    if t == torch.float16:
        return DVL_F16
    raise TypeError(f"unsupported embedding dtype {t}")


def _np_dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return DVL_F32
    if a.dtype == np.float16:
        return DVL_F16
    raise TypeError(f"unsupported host embedding dtype {a.dtype}")


def build_info() -> dict:
    """What the in-tree library was built from (csrc/build.py writes it next to the .so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dvl_build", os.path.join(_HERE, "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_info()


class Context:
    """One dvl_ctx bound to one CUDA device."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        rc = lib().dvl_create(ctypes.byref(self._h), int(device))
        if rc != DVL_OK:
            self._h = ctypes.c_void_p()
            raise DvlError(rc, f"dvl_create(device={device}) failed: an sm_100 (B200) device is required; "
                               "there is no CPU fallback")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().dvl_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != DVL_OK:
            raise DvlError(rc, lib().dvl_last_error(self._h).decode("utf-8", "replace"))

    @property
    def launch_count(self) -> int:
        return int(lib().dvl_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(lib().dvl_workspace_bytes(self._h))

    def set_profiling(self, enabled: bool):
        self._check(lib().dvl_set_profiling(self._h, 1 if enabled else 0))

    def get_profile(self):
        """{stage name: (milliseconds, kernel launches)} accumulated since the previous call."""
        ms = (ctypes.c_double * NUM_STAGES)()
        ln = (ctypes.c_uint64 * NUM_STAGES)()
        self._check(lib().dvl_get_profile(self._h, ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p)))
        return {STAGES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_STAGES)}

    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()

    # ------------------------------------------------------------------------------------
    def encode_labels(self, labels: torch.Tensor):
        """Device integer labels (int64 / int32 / uint8) -> (uint8 codes on the device, info dict with
        ``counts`` [16], ``n_classes`` = largest label + 1, ``out_of_range``, ``contiguous``)."""
        assert labels.is_cuda and labels.is_contiguous() and labels.dim() == 1
        code = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}.get(labels.dtype)
        if code is None:
            raise TypeError(f"unsupported label dtype {labels.dtype}")
        n = labels.numel()
        codes = torch.empty((n,), dtype=torch.uint8, device=labels.device)
        info = (ctypes.c_int64 * 19)()
        self._check(lib().dvl_encode_labels(self._h, self._ptr(labels), code, n, self._ptr(codes),
                                            ctypes.cast(info, ctypes.c_void_p), ctypes.c_void_p(self._stream())))
        return codes, dict(counts=[int(x) for x in info[:16]], n_classes=int(info[16]), out_of_range=bool(info[17]),
                           contiguous=bool(info[18]))

    def eval_ranking(self, gallery: torch.Tensor, queries: torch.Tensor, labels_u8: torch.Tensor, n_classes: int,
                     top_n: int, precision: int = PRECISION_FP32, flags: int = 0, want_per_query: bool = False,
                     ranked_cap: int = 0, want_scores: bool = False, out_per_query: Optional[torch.Tensor] = None,
                     out_ranked_len: Optional[torch.Tensor] = None):
        """gallery [N,D], queries [Q,D], labels uint8 [N]: all CUDA, contiguous.  Returns a dict with
        ``mean`` (4 python floats) and optionally ``per_query`` [Q,4] f64, ``ranked_len`` [Q] i32,
        ``ranked_idx`` [Q,ranked_cap] i32, ``scores`` [Q,N] f32 (device tensors).  ``out_per_query`` /
        ``out_ranked_len``: caller-owned output buffers (contiguous f64 [Q,4] / i32 [Q]); a loop that passes the
        same buffers every time keeps the call identical, which is what the CUDA-graph replay keys on."""
        assert gallery.is_cuda and queries.is_cuda and labels_u8.is_cuda
        assert gallery.is_contiguous() and queries.is_contiguous() and labels_u8.is_contiguous()
        assert labels_u8.dtype == torch.uint8
        n, d = gallery.shape
        q, d2 = queries.shape
        if d != d2:
            raise ValueError(f"embedding depth mismatch: gallery {d}, queries {d2}")
        if labels_u8.numel() != n:
            raise ValueError("labels and gallery differ in length")
        dev = gallery.device
        mean = (ctypes.c_double * 4)()
        per_query = torch.empty((q, 4), dtype=torch.float64, device=dev) if want_per_query else None
        ranked_len = torch.empty((q,), dtype=torch.int32, device=dev) if (ranked_cap > 0 or want_per_query) else None
        if out_per_query is not None:
            assert out_per_query.is_cuda and out_per_query.dtype == torch.float64 and out_per_query.is_contiguous()
            assert out_per_query.numel() == q * 4
            per_query = out_per_query
        if out_ranked_len is not None:
            assert out_ranked_len.is_cuda and out_ranked_len.dtype == torch.int32 and out_ranked_len.is_contiguous()
            assert out_ranked_len.numel() == q
            ranked_len = out_ranked_len
        ranked_idx = torch.full((q, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap > 0 else None
        scores = torch.empty((q, n), dtype=torch.float32, device=dev) if want_scores else None
        rc = lib().dvl_eval_ranking(self._h, self._ptr(gallery), _dtype_code(gallery.dtype), self._ptr(queries),
                                    _dtype_code(queries.dtype), self._ptr(labels_u8), n, d, q, int(n_classes),
                                    int(top_n), int(precision), int(flags), ctypes.cast(mean, ctypes.c_void_p),
                                    self._ptr(per_query), self._ptr(ranked_len), self._ptr(ranked_idx),
                                    int(ranked_cap), self._ptr(scores), ctypes.c_void_p(self._stream()))
        self._check(rc)
        return dict(mean=[float(x) for x in mean], per_query=per_query, ranked_len=ranked_len,
                    ranked_idx=ranked_idx, scores=scores)

    def eval_ranking_host(self, gallery, queries, labels_u8, n_classes: int, top_n: int,
                          precision: int = PRECISION_FP32, flags: int = 0, per_query_out: Optional[np.ndarray] = None):
        """HOST inputs (numpy arrays or CPU torch tensors, ideally pinned): the H2D copies happen inside the
        call, the gallery in row blocks that are scored as they land."""
        def host(x):
            if isinstance(x, torch.Tensor):
                assert not x.is_cuda and x.is_contiguous()
                return x.data_ptr(), (_dtype_code(x.dtype) if x.dtype != torch.uint8 else None), tuple(x.shape)
            assert x.flags["C_CONTIGUOUS"]
            return x.ctypes.data, (_np_dtype_code(x) if x.dtype != np.uint8 else None), tuple(x.shape)
        g_ptr, g_dt, (n, d) = host(gallery)
        t_ptr, t_dt, (q, d2) = host(queries)
        l_ptr, l_dt, l_shape = host(labels_u8)
        if d != d2:
            raise ValueError(f"embedding depth mismatch: gallery {d}, queries {d2}")
        if l_dt is not None or l_shape != (n,):
            raise ValueError("labels must be uint8 class codes, one per gallery row")
        mean = (ctypes.c_double * 4)()
        pq = ctypes.c_void_p(per_query_out.ctypes.data) if per_query_out is not None else ctypes.c_void_p()
        rc = lib().dvl_eval_ranking_host(self._h, ctypes.c_void_p(g_ptr), g_dt, ctypes.c_void_p(t_ptr), t_dt,
                                         ctypes.c_void_p(l_ptr), n, d, q, int(n_classes), int(top_n),
                                         int(precision), int(flags), ctypes.cast(mean, ctypes.c_void_p), pq,
                                         ctypes.c_void_p(self._stream()))
        self._check(rc)
        return [float(x) for x in mean]

    def rank_and_score(self, scores: torch.Tensor, labels_u8: torch.Tensor, n_classes: int, top_n: int,
                       ranked_cap: int = 0):
        assert scores.is_cuda and scores.dtype == torch.float32 and scores.stride(1) == 1
        q, n = scores.shape
        dev = scores.device
        mean = (ctypes.c_double * 4)()
        per_query = torch.empty((q, 4), dtype=torch.float64, device=dev)
        ranked_len = torch.empty((q,), dtype=torch.int32, device=dev)
        ranked_idx = torch.full((q, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap > 0 else None
        rc = lib().dvl_rank_and_score(self._h, self._ptr(scores), int(scores.stride(0)), self._ptr(labels_u8), n, q,
                                      int(n_classes), int(top_n), ctypes.cast(mean, ctypes.c_void_p),
                                      self._ptr(per_query), self._ptr(ranked_len), self._ptr(ranked_idx),
                                      int(ranked_cap), ctypes.c_void_p(self._stream()))
        self._check(rc)
        return dict(mean=[float(x) for x in mean], per_query=per_query, ranked_len=ranked_len, ranked_idx=ranked_idx)

    def similarity(self, gallery: torch.Tensor, queries: torch.Tensor, precision: int = PRECISION_FP32,
                   flags: int = 0) -> torch.Tensor:
        assert gallery.is_cuda and queries.is_cuda and gallery.is_contiguous() and queries.is_contiguous()
        n, d = gallery.shape
        q = queries.shape[0]
        out = torch.empty((q, n), dtype=torch.float32, device=gallery.device)
        rc = lib().dvl_similarity(self._h, self._ptr(gallery), _dtype_code(gallery.dtype), self._ptr(queries),
                                  _dtype_code(queries.dtype), n, d, q, int(precision), int(flags), self._ptr(out), n,
                                  ctypes.c_void_p(self._stream()))
        self._check(rc)
        return out


    # ---- gallery-sharded stages (see sharded.py) -------------------------------------------
    def shard_topk(self, gallery: torch.Tensor, queries: torch.Tensor, labels_u8: torch.Tensor, index_base: int,
                   n_classes: int, top_n: int, precision: int = PRECISION_FP32, flags: int = 0, outputs: bool = True):
        """``outputs=False``: no copies of tau / flags / label counts (the lean exchange reads them from the
        context with ``shard_stats``); returns ``(None, None, None, list_cap)``."""
        assert gallery.is_cuda and gallery.is_contiguous() and queries.is_contiguous() and labels_u8.dtype == torch.uint8
        n_local, d = gallery.shape
        q = queries.shape[0]
        tau, oflags, counts = self._shard_outputs(q, gallery.device, outputs)
        cap = ctypes.c_int32(0)
        rc = lib().dvl_shard_topk(self._h, self._ptr(gallery), _dtype_code(gallery.dtype), self._ptr(queries),
                                  _dtype_code(queries.dtype), self._ptr(labels_u8), n_local, int(index_base), d, q,
                                  int(n_classes), int(top_n), int(precision), int(flags), self._ptr(tau),
                                  self._ptr(oflags), self._ptr(counts), ctypes.cast(ctypes.byref(cap), ctypes.c_void_p),
                                  ctypes.c_void_p(self._stream()))
        self._check(rc)
        return tau, oflags, counts, int(cap.value)

    @staticmethod
    def _shard_outputs(q: int, dev, outputs: bool):
        if not outputs:
            return None, None, None
        return (torch.empty((q,), dtype=torch.float32, device=dev), torch.empty((q,), dtype=torch.int32, device=dev),
                torch.empty((MAX_CLASSES + 1,), dtype=torch.int64, device=dev))

    def shard_topk_host(self, gallery: torch.Tensor, queries: torch.Tensor, labels_u8: torch.Tensor, index_base: int,
                        n_classes: int, top_n: int, precision: int = PRECISION_FP32, flags: int = 0,
                        outputs: bool = True):
        """``shard_topk`` with HOST tensors (pinned for an asynchronous, overlapped upload): the gallery shard
        travels in row blocks on the context's copy stream while the blocks that have landed are already
        being scored and filtered."""
        assert not gallery.is_cuda and not queries.is_cuda and not labels_u8.is_cuda
        assert gallery.is_contiguous() and queries.is_contiguous() and labels_u8.is_contiguous()
        assert labels_u8.dtype == torch.uint8
        n_local, d = gallery.shape
        q = queries.shape[0]
        tau, oflags, counts = self._shard_outputs(q, torch.device("cuda", self.device), outputs)
        cap = ctypes.c_int32(0)
        rc = lib().dvl_shard_topk_host(self._h, self._ptr(gallery), _dtype_code(gallery.dtype), self._ptr(queries),
                                       _dtype_code(queries.dtype), self._ptr(labels_u8), n_local, int(index_base), d, q,
                                       int(n_classes), int(top_n), int(precision), int(flags), self._ptr(tau),
                                       self._ptr(oflags), self._ptr(counts),
                                       ctypes.cast(ctypes.byref(cap), ctypes.c_void_p), None, None, None,
                                       ctypes.c_void_p(self._stream()))
        self._check(rc)
        return tau, oflags, counts, int(cap.value)

    # the lean exchange (caller-owned persistent buffers, no allocation, one host synchronisation)
    def shard_stats(self, out_stats: torch.Tensor):
        """This rank's statistics record [tau | flags | label counts] as 2q + 34 int32 words."""
        q = (out_stats.numel() - 2 * (MAX_CLASSES + 1)) // 2
        assert out_stats.is_cuda and out_stats.dtype == torch.int32 and out_stats.is_contiguous()
        self._check(lib().dvl_shard_stats(self._h, self._ptr(out_stats), q, ctypes.c_void_p(self._stream())))

    def shard_reduce(self, stats_all: torch.Tensor, q: int):
        """[world, 2q + 34] gathered records -> tau bound / flags / global label counts kept in the context."""
        assert stats_all.is_cuda and stats_all.dtype == torch.int32 and stats_all.is_contiguous()
        self._check(lib().dvl_shard_reduce(self._h, self._ptr(stats_all), stats_all.shape[0], int(q),
                                           ctypes.c_void_p(self._stream())))

    def shard_trim_pack(self, out_packed: torch.Tensor, out_need: torch.Tensor):
        """Local lists cut at the global bound, packed into out_packed [q, len]; out_need [1] i32 = longest."""
        q, length = out_packed.shape
        assert out_packed.is_contiguous() and out_packed.dtype == torch.int64 and out_need.dtype == torch.int32
        self._check(lib().dvl_shard_trim_pack(self._h, q, length, self._ptr(out_packed), self._ptr(out_need),
                                              ctypes.c_void_p(self._stream())))

    def merge_topk_into(self, gathered: torch.Tensor, q0: int, qn: int, n_global: int, n_classes: int, top_n: int,
                        per_query: torch.Tensor, ranked_len: torch.Tensor, flags: torch.Tensor,
                        ranked_idx: Optional[torch.Tensor] = None):
        """``merge_topk`` writing into caller-owned buffers, against the label counts ``shard_reduce`` summed."""
        world, q_total, length = gathered.shape
        assert gathered.is_contiguous() and gathered.dtype == torch.int64
        assert per_query.dtype == torch.float64 and ranked_len.dtype == torch.int32 and flags.dtype == torch.int32
        cap = ranked_idx.shape[1] if ranked_idx is not None else 0
        rc = lib().dvl_merge_topk(self._h, self._ptr(gathered), world, q_total, length, int(q0), int(qn), int(n_global),
                                  None, int(n_classes), int(top_n), self._ptr(per_query), self._ptr(ranked_len),
                                  self._ptr(ranked_idx), int(cap), self._ptr(flags), ctypes.c_void_p(self._stream()))
        self._check(rc)

    def shard_finish(self, records_all: torch.Tensor, block: int, q: int):
        """[world, rec_bytes] gathered result records -> (four means, (any_flag, bad_label, need_max)); the
        exchange's one host synchronisation."""
        assert records_all.is_cuda and records_all.dtype == torch.uint8 and records_all.is_contiguous()
        world, rec_bytes = records_all.shape
        mean = (ctypes.c_double * 4)()
        status = (ctypes.c_int64 * 3)()
        self._check(lib().dvl_shard_finish(self._h, self._ptr(records_all), world, int(block), rec_bytes, int(q),
                                           ctypes.cast(mean, ctypes.c_void_p), ctypes.cast(status, ctypes.c_void_p),
                                           ctypes.c_void_p(self._stream())))
        return [float(x) for x in mean], tuple(int(x) for x in status)

    def shard_trim(self, tau_bound: torch.Tensor) -> torch.Tensor:
        q = tau_bound.numel()
        cnt = torch.empty((q,), dtype=torch.int32, device=tau_bound.device)
        self._check(lib().dvl_shard_trim(self._h, self._ptr(tau_bound), q, self._ptr(cnt),
                                         ctypes.c_void_p(self._stream())))
        return cnt

    def shard_pack(self, cnt: torch.Tensor, length: int) -> torch.Tensor:
        q = cnt.numel()
        out = torch.empty((q, length), dtype=torch.int64, device=cnt.device)    # u64 entries, stored as i64
        self._check(lib().dvl_shard_pack(self._h, self._ptr(cnt), q, int(length), self._ptr(out),
                                         ctypes.c_void_p(self._stream())))
        return out

    def merge_topk(self, gathered: torch.Tensor, q0: int, qn: int, n_global: int, class_counts: torch.Tensor,
                   n_classes: int, top_n: int, ranked_cap: int = 0):
        world, q_total, length = gathered.shape
        assert gathered.is_contiguous() and gathered.dtype == torch.int64 and class_counts.dtype == torch.int64
        dev = gathered.device
        per_query = torch.empty((qn, 4), dtype=torch.float64, device=dev)
        ranked_len = torch.empty((qn,), dtype=torch.int32, device=dev)
        ranked_idx = torch.full((qn, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap > 0 else None
        oflags = torch.empty((qn,), dtype=torch.int32, device=dev)
        rc = lib().dvl_merge_topk(self._h, self._ptr(gathered), world, q_total, length, int(q0), int(qn), int(n_global),
                                  self._ptr(class_counts), int(n_classes), int(top_n), self._ptr(per_query),
                                  self._ptr(ranked_len), self._ptr(ranked_idx), int(ranked_cap), self._ptr(oflags),
                                  ctypes.c_void_p(self._stream()))
        self._check(rc)
        return dict(per_query=per_query, ranked_len=ranked_len, ranked_idx=ranked_idx, flags=oflags)

    def mean(self, per_query: torch.Tensor):
        assert per_query.is_cuda and per_query.dtype == torch.float64 and per_query.is_contiguous()
        mean = (ctypes.c_double * 4)()
        self._check(lib().dvl_mean(self._h, self._ptr(per_query), per_query.shape[0], ctypes.cast(mean, ctypes.c_void_p),
                                   ctypes.c_void_p(self._stream())))
        return [float(x) for x in mean]


_contexts = {}


def context(device: Optional[int] = None) -> Context:
    """Process-wide context per device."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
