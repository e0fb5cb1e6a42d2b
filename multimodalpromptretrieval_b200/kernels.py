"""torch-tensor front end of the C ABI: pointers, sizes and the current CUDA stream go down, nothing else.

torch is plumbing here (device memory + streams); all arithmetic happens in ``libmpr_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _native

_handles = {}

# NVTX ranges around bank build / retrieval step / tokenisation / shard I/O (MPR_NVTX=1): named spans for nsys / ncu
# timelines.  Off by default — the step's host path is a few tens of microseconds and two extra calls would show.
_NVTX = os.environ.get("MPR_NVTX", "") == "1"


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def handle(device: Optional[int] = None) -> _native.Handle:
    h = _handles.get(device) if device is not None else None      # hot path: one dict lookup
    if h is not None:
        return h
    if not torch.cuda.is_available():
        raise _native.NativeError("no CUDA device: this package runs on B200 (sm_100a) only and has no fallback path")
    if device is None:
        device = torch.cuda.current_device()
    h = _handles.get(device)
    if h is None:
        h = _native.Handle(device)
        _handles[device] = h
    return h


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(device=None) -> C.c_void_p:
    """The caller's current stream ON THE TENSORS' DEVICE (not on whatever device happens to be current); the library
    itself switches to its handle's device for the launch."""
    if isinstance(device, torch.device) and device.index is not None:
        # the raw handle of the current stream, without building a torch.cuda.Stream object (~10 us per step otherwise)
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(device.index))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_DTYPES = {torch.float32: _native.SRC_F32, torch.float16: _native.SRC_F16, torch.bfloat16: _native.SRC_BF16}


def bank_build(src0: torch.Tensor, src1: Optional[torch.Tensor] = None, normalise: bool = False,
               out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernel 1: ``[n,d0] (‖ [n,d1])`` fp32/fp16/bf16 → bf16 ``[n,d]`` rows + fp32 ``bias[n] = -½‖row‖²``."""
    h = handle(src0.device.index)
    assert src0.is_cuda and src0.dim() == 2 and src0.is_contiguous()
    n, d0 = src0.shape
    d1 = 0
    if src1 is not None:
        assert src1.is_cuda and src1.is_contiguous() and src1.shape[0] == n and src1.dtype == src0.dtype
        d1 = src1.shape[1]
    if out is None:
        out = torch.empty((n, d0 + d1), dtype=torch.bfloat16, device=src0.device)
    if bias is None:
        bias = torch.empty((n,), dtype=torch.float32, device=src0.device)
    assert out.is_contiguous() and out.dtype == torch.bfloat16 and tuple(out.shape) == (n, d0 + d1)
    rc = h.lib.mpr_bank_build(h.ptr, _ptr(src0), d0, _ptr(src1), d1, _DTYPES[src0.dtype], n, int(bool(normalise)),
                              _ptr(out), _ptr(bias), _stream(src0.device))
    h.check(rc, "mpr_bank_build")
    return out, bias


def search_plan(b: int, n_local: int, d: int, kk: int, device: Optional[int] = None) -> dict:
    h = handle(device)
    vals = [C.c_int(0) for _ in range(5)]
    h.check(h.lib.mpr_search_plan(h.ptr, b, n_local, d, kk, *[C.byref(v) for v in vals]), "mpr_search_plan")
    keys = ["n_ctas", "n_splits", "n_qtiles", "n_stages", "smem_bytes"]
    return {k: v.value for k, v in zip(keys, vals)}


def search_workspace_bytes(b: int, n_local: int, d: int, kk: int, device: Optional[int] = None) -> int:
    h = handle(device)
    return int(h.lib.mpr_search_workspace_bytes(h.ptr, b, n_local, d, kk))


def new_workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """A fresh search workspace.  The caching allocator may hand back an address the library has seen before (with
    control words it believes to be zero), so the library is told to forget that address."""
    ws = torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=device)
    h = handle(device.index)
    h.check(h.lib.mpr_workspace_invalidate(h.ptr, _ptr(ws)), "mpr_workspace_invalidate")
    return ws


def search_topk(q: torch.Tensor, bank: torch.Tensor, bias: torch.Tensor, kk: int, idx_base: int = 0,
                workspace: Optional[torch.Tensor] = None, out_keys: Optional[torch.Tensor] = None,
                out_score: Optional[torch.Tensor] = None, out_idx: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Kernel 2 (+4): top-``kk`` rows of this shard for every query → (keys u64-as-int64, score f32, idx i32)."""
    h = handle(q.device.index)
    assert q.dtype == torch.bfloat16 and bank.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert q.is_contiguous() and bank.is_contiguous() and bias.is_contiguous()
    b, d = q.shape
    n_local = bank.shape[0]
    assert bank.shape[1] == d and bias.shape[0] == n_local
    dev = q.device
    if out_keys is None:
        out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev)
    if out_score is None:
        out_score = torch.empty((b, kk), dtype=torch.float32, device=dev)
    if out_idx is None:
        out_idx = torch.empty((b, kk), dtype=torch.int32, device=dev)
    need = search_workspace_bytes(b, n_local, d, kk, dev.index)
    if need == 0 and b > 0:
        # let the library produce the precise validation error
        rc = h.lib.mpr_search_plan(h.ptr, b, n_local, d, kk, None, None, None, None, None)
        h.check(rc, "mpr_search_plan")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = new_workspace(need, dev)
    rc = h.lib.mpr_search_topk(h.ptr, _ptr(q), b, _ptr(bank), _ptr(bias), n_local, idx_base, d, kk, _ptr(out_keys),
                               _ptr(out_score), _ptr(out_idx), _ptr(workspace),
                               workspace.numel() * workspace.element_size(), _stream(dev))
    h.check(rc, "mpr_search_topk")
    return out_keys, out_score, out_idx


def search_fused_supported(d: int, device: Optional[int] = None) -> bool:
    h = handle(device)
    return bool(h.lib.mpr_search_fused_supported(h.ptr, int(d)))


def search_topk_fused(src0: torch.Tensor, src1: Optional[torch.Tensor], bank: torch.Tensor, bias: torch.Tensor, kk: int,
                      normalise: bool = False, idx_base: int = 0, workspace: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Kernel 2 with the query concat / normalise / bf16 cast fused into its q-tile load (D <= 512).
    Returns (keys, score, idx, q_bias)."""
    h = handle(src0.device.index)
    assert src0.is_cuda and src0.dim() == 2 and src0.is_contiguous() and bank.dtype == torch.bfloat16
    b, d0 = src0.shape
    d1 = 0
    if src1 is not None:
        assert src1.is_contiguous() and src1.shape[0] == b and src1.dtype == src0.dtype
        d1 = src1.shape[1]
    d = d0 + d1
    n_local = bank.shape[0]
    assert bank.shape[1] == d and bias.shape[0] == n_local
    dev = src0.device
    out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev)
    out_score = torch.empty((b, kk), dtype=torch.float32, device=dev)
    out_idx = torch.empty((b, kk), dtype=torch.int32, device=dev)
    q_bias = torch.empty((b,), dtype=torch.float32, device=dev)
    need = search_workspace_bytes(b, n_local, d, kk, dev.index)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = new_workspace(need, dev)
    rc = h.lib.mpr_search_topk_fused(h.ptr, _ptr(src0), d0, _ptr(src1), d1, _DTYPES[src0.dtype], int(bool(normalise)), b,
                                     _ptr(bank), _ptr(bias), n_local, idx_base, kk, _ptr(out_keys), _ptr(out_score),
                                     _ptr(out_idx), _ptr(q_bias), _ptr(workspace),
                                     workspace.numel() * workspace.element_size(), _stream(dev))
    h.check(rc, "mpr_search_topk_fused")
    return out_keys, out_score, out_idx, q_bias


def merge_topk(keys: torch.Tensor, out_keys: Optional[torch.Tensor] = None, out_score: Optional[torch.Tensor] = None,
               out_idx: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Kernel 4: ``keys [n_lists, b, kk]`` (each list sorted) → global top-kk."""
    h = handle(keys.device.index)
    assert keys.dtype == torch.int64 and keys.dim() == 3 and keys.is_contiguous()
    n_lists, b, kk = keys.shape
    dev = keys.device
    if out_keys is None:
        out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev)
    if out_score is None:
        out_score = torch.empty((b, kk), dtype=torch.float32, device=dev)
    if out_idx is None:
        out_idx = torch.empty((b, kk), dtype=torch.int32, device=dev)
    rc = h.lib.mpr_merge_topk(h.ptr, _ptr(keys), n_lists, b, kk, _ptr(out_keys), _ptr(out_score), _ptr(out_idx),
                              _stream(dev))
    h.check(rc, "mpr_merge_topk")
    return out_keys, out_score, out_idx


def prompt_gather(idx: torch.Tensor, skip: int, answer_id: torch.Tensor, bucket_lut: torch.Tensor,
                  prefix_ids: torch.Tensor, prefix_off: torch.Tensor, seg_ids: torch.Tensor, seg_off: torch.Tensor,
                  use_quantifier: bool, pad_id: int, eos_id: int, max_len: int, out_stride: int) -> dict:
    """Kernel 3: retrieved rows → vote → bucket → ``input_ids`` / ``attention_mask`` (int64, padded to out_stride)."""
    h = handle(idx.device.index)
    assert idx.dtype == torch.int32 and idx.is_contiguous() and idx.dim() == 2
    for t in (answer_id, prefix_ids, prefix_off, seg_ids, seg_off):
        assert t.dtype == torch.int32 and t.is_cuda and t.is_contiguous()
    assert bucket_lut.dtype == torch.uint8 and bucket_lut.is_cuda
    b, kk = idx.shape
    k = kk - skip
    assert bucket_lut.numel() == (k + 1) * (k + 1)
    dev = idx.device
    out = {
        "input_ids": torch.empty((b, out_stride), dtype=torch.int64, device=dev),
        "attention_mask": torch.empty((b, out_stride), dtype=torch.int64, device=dev),
        "length": torch.empty((b,), dtype=torch.int32, device=dev),
        "majority_answer": torch.empty((b,), dtype=torch.int32, device=dev),
        "majority_count": torch.empty((b,), dtype=torch.int32, device=dev),
        "bucket": torch.empty((b,), dtype=torch.int32, device=dev),
        "answer_ids": torch.empty((b, k), dtype=torch.int32, device=dev),
    }
    rc = h.lib.mpr_prompt_gather(
        h.ptr, _ptr(idx), b, kk, skip, _ptr(answer_id), _ptr(bucket_lut), _ptr(prefix_ids), _ptr(prefix_off),
        _ptr(seg_ids), _ptr(seg_off), int(bool(use_quantifier)), pad_id, eos_id, max_len, out_stride,
        _ptr(out["input_ids"]), _ptr(out["attention_mask"]), _ptr(out["length"]), _ptr(out["majority_answer"]),
        _ptr(out["majority_count"]), _ptr(out["bucket"]), _ptr(out["answer_ids"]), _stream(dev))
    h.check(rc, "mpr_prompt_gather")
    return out


def debug_scores(q: torch.Tensor, bank: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Full ``[b, n]`` score matrix through the same tcgen05 pipeline as :func:`search_topk` (tests only)."""
    h = handle(q.device.index)
    b, d = q.shape
    n_local = bank.shape[0]
    scores = torch.full((b, n_local), float("nan"), dtype=torch.float32, device=q.device)
    need = max(16, search_workspace_bytes(b, n_local, d, 1, q.device.index))
    ws = new_workspace(need, q.device)
    rc = h.lib.mpr_debug_scores(h.ptr, _ptr(q), b, _ptr(bank), _ptr(bias), n_local, d, _ptr(scores), _ptr(ws), need,
                                _stream(q.device))
    h.check(rc, "mpr_debug_scores")
    return scores


def profile_begin(max_launches: int, device: Optional[int] = None) -> None:
    h = handle(device)
    h.check(h.lib.mpr_profile_begin(h.ptr, max_launches), "mpr_profile_begin")


def profile_end(device: Optional[int] = None) -> Tuple[float, int]:
    """(summed scan-kernel device time in ms, launches recorded) since :func:`profile_begin`."""
    h = handle(device)
    ms, n = C.c_float(0), C.c_int(0)
    h.check(h.lib.mpr_profile_end(h.ptr, C.byref(ms), C.byref(n)), "mpr_profile_end")
    return ms.value, n.value


def profile_launches(n: int, device: Optional[int] = None) -> list:
    """Per-launch scan-kernel times (ms) of the last profile_begin/profile_end pair."""
    h = handle(device)
    out = []
    for i in range(n):
        ms = C.c_float(0)
        h.check(h.lib.mpr_profile_launch_ms(h.ptr, i, C.byref(ms)), "mpr_profile_launch_ms")
        out.append(ms.value)
    return out


def exchange_bytes(world: int, cap: int) -> int:
    return int(_native.load().mpr_exchange_bytes(world, cap))


def retrieve(args: "_native.RetrieveArgs", device: torch.device, io: Optional["_native.HostIO"] = None) -> None:
    """One retrieval step (``mpr_retrieve`` / ``mpr_retrieve_host``) from a pre-filled argument block: the hot call of
    :class:`~multimodalpromptretrieval_b200.bank.RetrievalBank` — one ctypes call, normally one kernel launch."""
    h = handle(device.index)
    if io is None:
        h.check(h.lib.mpr_retrieve(h.ptr, C.byref(args), _stream(device)), "mpr_retrieve")
    else:
        h.check(h.lib.mpr_retrieve_host(h.ptr, C.byref(args), C.byref(io), _stream(device)), "mpr_retrieve_host")


def retrieve_join(device: torch.device) -> None:
    """Makes the current stream wait for every deferred finish queued so far (``RetrieveArgs.defer_finish``)."""
    h = handle(device.index)
    h.check(h.lib.mpr_retrieve_join(h.ptr, _stream(device)), "mpr_retrieve_join")


def debug_counters(device: Optional[int] = None) -> dict:
    """Scan-kernel event counters since the last call (all zero unless MPR_DEBUG_COUNTERS=1 was set at handle creation)."""
    h = handle(device)
    out = (C.c_uint64 * 8)()
    h.check(h.lib.mpr_debug_counters(h.ptr, out), "mpr_debug_counters")
    names = ["candidates", "flushes", "slow_groups", "replacements", "warp_tiles", "bound_refreshes"]
    return {n: int(out[i]) for i, n in enumerate(names)}


def debug_timeline(n_ctas: int, device: Optional[int] = None):
    """Per-CTA event timestamps (ns, relative to the earliest CTA entry) of the last scan launch; see mpr_b200.h."""
    import numpy as np
    h = handle(device)
    out = (C.c_uint64 * (24 * n_ctas))()
    h.check(h.lib.mpr_debug_timeline(h.ptr, out, n_ctas), "mpr_debug_timeline")
    a = np.frombuffer(out, dtype=np.uint64).reshape(n_ctas, 24).astype(np.int64)
    t0 = a[:, 0][a[:, 0] > 0].min() if (a[:, 0] > 0).any() else 0
    return np.where(a > 0, a - t0, -1)


def debug_launch_ring(device: Optional[int] = None):
    """(entry_ns, exit_ns) of the most recent scan launches, oldest first (MPR_DEBUG_COUNTERS=1|2); clears the ring."""
    h = handle(device)
    out = (C.c_uint64 * 128)()
    seq = C.c_uint(0)
    h.check(h.lib.mpr_debug_launch_ring(h.ptr, out, C.byref(seq)), "mpr_debug_launch_ring")
    res = []
    for s_ in range(max(0, seq.value - 64), seq.value):
        first, last = int(out[2 * (s_ % 64)]), int(out[2 * (s_ % 64) + 1])
        if first and last:
            res.append(((1 << 63) - first, last))
    return res


def last_launch_count(device: Optional[int] = None) -> int:
    h = handle(device)
    return int(h.lib.mpr_last_launch_count(h.ptr))


def set_exchange_timeout(seconds: float, device: Optional[int] = None) -> None:
    h = handle(device)
    h.check(h.lib.mpr_set_exchange_timeout(h.ptr, float(seconds)), "mpr_set_exchange_timeout")


def embed_prompt(input_ids: torch.Tensor, attention_mask: torch.Tensor, table: torch.Tensor,
                 image_tokens: Optional[torch.Tensor] = None, length: Optional[int] = None,
                 mask_dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernel 5: ``input_ids [b, stride]`` (first ``length`` columns) -> ``[b, n_image + length, hidden]`` embeddings
    with the image tokens prepended, and the concatenated attention mask (float32 as in the reference, or int64)."""
    h = handle(input_ids.device.index)
    assert input_ids.dtype == torch.int64 and attention_mask.dtype == torch.int64 and input_ids.dim() == 2
    assert input_ids.stride(1) == 1 and attention_mask.stride(1) == 1 and input_ids.stride(0) == attention_mask.stride(0)
    assert table.is_cuda and table.is_contiguous() and table.dim() == 2 and table.dtype in _DTYPES
    b, stride = input_ids.shape[0], input_ids.stride(0)
    length = input_ids.shape[1] if length is None else int(length)
    vocab, hidden = table.shape
    n_image = 0
    if image_tokens is not None:
        assert image_tokens.is_contiguous() and image_tokens.dtype == table.dtype and image_tokens.shape[0] == b and \
            image_tokens.shape[2] == hidden
        n_image = image_tokens.shape[1]
    assert mask_dtype in (torch.float32, torch.int64)
    dev = input_ids.device
    out = torch.empty((b, n_image + length, hidden), dtype=table.dtype, device=dev)
    mask = torch.empty((b, n_image + length), dtype=mask_dtype, device=dev)
    rc = h.lib.mpr_embed_prompt(h.ptr, _ptr(input_ids), _ptr(attention_mask), b, length, stride, _ptr(table),
                                _DTYPES[table.dtype], vocab, hidden, _ptr(image_tokens), n_image, _ptr(out), _ptr(mask),
                                int(mask_dtype == torch.float32), _stream(dev))
    h.check(rc, "mpr_embed_prompt")
    return out, mask
