"""torch-tensor front end of the C ABI: pointers, sizes and the current CUDA stream go down, nothing else.

torch is plumbing here (device memory + streams); all arithmetic happens in ``libmpr_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _native

_handles = {}


def handle(device: Optional[int] = None) -> _native.Handle:
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


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


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
                              _ptr(out), _ptr(bias), _stream())
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
        workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=dev)
    rc = h.lib.mpr_search_topk(h.ptr, _ptr(q), b, _ptr(bank), _ptr(bias), n_local, idx_base, d, kk, _ptr(out_keys),
                               _ptr(out_score), _ptr(out_idx), _ptr(workspace),
                               workspace.numel() * workspace.element_size(), _stream())
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
        workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=dev)
    rc = h.lib.mpr_search_topk_fused(h.ptr, _ptr(src0), d0, _ptr(src1), d1, _DTYPES[src0.dtype], int(bool(normalise)), b,
                                     _ptr(bank), _ptr(bias), n_local, idx_base, kk, _ptr(out_keys), _ptr(out_score),
                                     _ptr(out_idx), _ptr(q_bias), _ptr(workspace),
                                     workspace.numel() * workspace.element_size(), _stream())
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
                              _stream())
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
        _ptr(out["majority_count"]), _ptr(out["bucket"]), _ptr(out["answer_ids"]), _stream())
    h.check(rc, "mpr_prompt_gather")
    return out


def debug_scores(q: torch.Tensor, bank: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Full ``[b, n]`` score matrix through the same tcgen05 pipeline as :func:`search_topk` (tests only)."""
    h = handle(q.device.index)
    b, d = q.shape
    n_local = bank.shape[0]
    scores = torch.full((b, n_local), float("nan"), dtype=torch.float32, device=q.device)
    need = max(16, search_workspace_bytes(b, n_local, d, 1, q.device.index))
    ws = torch.empty((need,), dtype=torch.uint8, device=q.device)
    rc = h.lib.mpr_debug_scores(h.ptr, _ptr(q), b, _ptr(bank), _ptr(bias), n_local, d, _ptr(scores), _ptr(ws), need,
                                _stream())
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


def exchange_push(keys: torch.Tensor, rank: int, peer_ptrs, cap: int) -> None:
    """P2P push of this rank's candidate keys ``[b, kk]`` into every rank's exchange buffer (``peer_ptrs[r]`` = device
    address of rank r's buffer as mapped on this device)."""
    h = handle(keys.device.index)
    assert keys.dtype == torch.int64 and keys.dim() == 2 and keys.is_contiguous()
    b, kk = keys.shape
    world = len(peer_ptrs)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
    h.check(h.lib.mpr_exchange_push(h.ptr, _ptr(keys), b, kk, rank, world, arr, cap, _stream()), "mpr_exchange_push")


def exchange_merge(my_buf: torch.Tensor, world: int, cap: int, b: int, kk: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Wait for all ``world`` deliveries into ``my_buf`` and merge them into the global top-kk."""
    h = handle(my_buf.device.index)
    dev = my_buf.device
    out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev)
    out_score = torch.empty((b, kk), dtype=torch.float32, device=dev)
    out_idx = torch.empty((b, kk), dtype=torch.int32, device=dev)
    rc = h.lib.mpr_exchange_merge(h.ptr, _ptr(my_buf), world, cap, b, kk, _ptr(out_keys), _ptr(out_score),
                                  _ptr(out_idx), _stream())
    h.check(rc, "mpr_exchange_merge")
    return out_keys, out_score, out_idx
