"""ctypes binding of the C ABI declared in ``include/mpr_b200.h``.

The shared library is the product: if it is missing, or the device is not a B200-class (sm_100) GPU, every entry
point raises — there is no eager / CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmpr_b200.so")

MPR_MAX_KK = 32
SRC_F32, SRC_F16, SRC_BF16 = 0, 1, 2

ABI_VERSION = 2
STATUS_XCHG_TIMEOUT = 201

EXPORTS = [
    "mpr_abi_version", "mpr_create", "mpr_destroy", "mpr_last_error", "mpr_device_error", "mpr_bank_build",
    "mpr_search_workspace_bytes", "mpr_search_topk", "mpr_merge_topk", "mpr_prompt_gather", "mpr_debug_scores",
    "mpr_search_plan", "mpr_plan_host", "mpr_profile_begin", "mpr_profile_end", "mpr_profile_launch_ms", "mpr_exchange_bytes",
    "mpr_search_fused_supported", "mpr_search_topk_fused", "mpr_retrieve", "mpr_retrieve_host",
    "mpr_set_exchange_timeout", "mpr_embed_prompt", "mpr_last_launch_count", "mpr_debug_counters", "mpr_debug_timeline", "mpr_debug_launch_ring", "mpr_retrieve_join", "mpr_workspace_invalidate",
    "mpr_token_cache_create", "mpr_token_cache_destroy", "mpr_token_cache_size", "mpr_token_cache_clear",
    "mpr_token_cache_put", "mpr_token_cache_assemble",
]


class RetrieveArgs(C.Structure):
    """``mpr_retrieve_args`` of include/mpr_b200.h, field for field."""
    _fields_ = [
        ("q0", C.c_void_p), ("q1", C.c_void_p), ("d0", C.c_int), ("d1", C.c_int), ("q_dtype", C.c_int),
        ("normalise", C.c_int), ("q_bf16", C.c_void_p), ("q_scratch", C.c_void_p), ("b", C.c_int),
        ("bank", C.c_void_p), ("bias", C.c_void_p), ("n_local", C.c_int64), ("idx_base", C.c_int64),
        ("d", C.c_int), ("kk", C.c_int),
        ("out_keys", C.c_void_p), ("out_score", C.c_void_p), ("out_idx", C.c_void_p), ("out_q_bias", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("rank", C.c_int), ("world", C.c_int), ("xchg_cap", C.c_int), ("peer_bufs", C.POINTER(C.c_void_p)),
        ("skip", C.c_int), ("answer_id", C.c_void_p), ("bucket_lut", C.c_void_p), ("prefix_ids", C.c_void_p),
        ("prefix_off", C.c_void_p), ("seg_ids", C.c_void_p), ("seg_off", C.c_void_p),
        ("use_quantifier", C.c_int), ("pad_id", C.c_int), ("eos_id", C.c_int), ("max_len", C.c_int),
        ("out_stride", C.c_int),
        ("input_ids", C.c_void_p), ("attention_mask", C.c_void_p), ("out_len", C.c_void_p), ("maj_answer", C.c_void_p),
        ("maj_count", C.c_void_p), ("bucket", C.c_void_p), ("ret_answer", C.c_void_p), ("status", C.c_void_p),
        ("defer_finish", C.c_int),
    ]


class HostIO(C.Structure):
    """``mpr_host_io`` of include/mpr_b200.h."""
    _fields_ = [
        ("h_q0", C.c_void_p), ("h_q1", C.c_void_p), ("h_prefix_ids", C.c_void_p), ("n_prefix_ids", C.c_int),
        ("h_prefix_off", C.c_void_p), ("d_out", C.c_void_p), ("h_out", C.c_void_p), ("out_bytes", C.c_size_t),
        ("sync", C.c_int), ("stream_in", C.c_void_p), ("stream_out", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None


class NativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libmpr_b200.so (built by ``__graft_entry__.build()`` / ``csrc/build.py``) and declare signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -m multimodalpromptretrieval_b200.csrc.build` "
            "(this package has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.mpr_abi_version.restype = i32
    lib.mpr_abi_version.argtypes = []
    lib.mpr_create.restype = i32
    lib.mpr_create.argtypes = [i32, C.POINTER(vp)]
    lib.mpr_destroy.restype = i32
    lib.mpr_destroy.argtypes = [vp]
    lib.mpr_last_error.restype = C.c_char_p
    lib.mpr_last_error.argtypes = [vp]
    lib.mpr_device_error.restype = i32
    lib.mpr_device_error.argtypes = [vp, C.POINTER(i32)]
    lib.mpr_bank_build.restype = i32
    lib.mpr_bank_build.argtypes = [vp, vp, i32, vp, i32, i32, i64, i32, vp, vp, vp]
    lib.mpr_search_workspace_bytes.restype = sz
    lib.mpr_search_workspace_bytes.argtypes = [vp, i32, i64, i32, i32]
    lib.mpr_search_topk.restype = i32
    lib.mpr_search_topk.argtypes = [vp, vp, i32, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.mpr_merge_topk.restype = i32
    lib.mpr_merge_topk.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.mpr_prompt_gather.restype = i32
    lib.mpr_prompt_gather.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                      vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mpr_debug_scores.restype = i32
    lib.mpr_debug_scores.argtypes = [vp, vp, i32, vp, vp, i64, i32, vp, vp, sz, vp]
    lib.mpr_plan_host.restype = i32
    lib.mpr_plan_host.argtypes = [i32, i32, i64, i32, i32, C.POINTER(i32)]
    lib.mpr_search_plan.restype = i32
    lib.mpr_search_plan.argtypes = [vp, i32, i64, i32, i32] + [C.POINTER(i32)] * 5
    lib.mpr_profile_begin.restype = i32
    lib.mpr_profile_begin.argtypes = [vp, i32]
    lib.mpr_profile_end.restype = i32
    lib.mpr_profile_end.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    lib.mpr_search_fused_supported.restype = i32
    lib.mpr_search_fused_supported.argtypes = [vp, i32]
    lib.mpr_search_topk_fused.restype = i32
    lib.mpr_search_topk_fused.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.mpr_exchange_bytes.restype = sz
    lib.mpr_exchange_bytes.argtypes = [i32, i32]
    lib.mpr_retrieve.restype = i32
    lib.mpr_retrieve.argtypes = [vp, C.POINTER(RetrieveArgs), vp]
    lib.mpr_retrieve_host.restype = i32
    lib.mpr_retrieve_host.argtypes = [vp, C.POINTER(RetrieveArgs), C.POINTER(HostIO), vp]
    lib.mpr_set_exchange_timeout.restype = i32
    lib.mpr_set_exchange_timeout.argtypes = [vp, C.c_double]
    lib.mpr_embed_prompt.restype = i32
    lib.mpr_embed_prompt.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, i32, vp, vp, i32, vp]
    lib.mpr_token_cache_create.restype = i32
    lib.mpr_token_cache_create.argtypes = [C.POINTER(vp)]
    lib.mpr_token_cache_destroy.restype = i32
    lib.mpr_token_cache_destroy.argtypes = [vp]
    lib.mpr_token_cache_size.restype = i64
    lib.mpr_token_cache_size.argtypes = [vp]
    lib.mpr_token_cache_clear.restype = i32
    lib.mpr_token_cache_clear.argtypes = [vp]
    lib.mpr_token_cache_put.restype = i32
    lib.mpr_token_cache_put.argtypes = [vp, i32, C.c_char_p, vp, vp, vp]
    lib.mpr_token_cache_assemble.restype = i32
    lib.mpr_token_cache_assemble.argtypes = [vp, i32, C.c_char_p, vp, vp, vp, vp, vp, i64, vp, vp, i32, C.POINTER(i32),
                                             C.POINTER(i32)]
    lib.mpr_workspace_invalidate.restype = i32
    lib.mpr_workspace_invalidate.argtypes = [vp, vp]
    lib.mpr_retrieve_join.restype = i32
    lib.mpr_retrieve_join.argtypes = [vp, vp]
    lib.mpr_debug_launch_ring.restype = i32
    lib.mpr_debug_launch_ring.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint)]
    lib.mpr_debug_timeline.restype = i32
    lib.mpr_debug_timeline.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    lib.mpr_debug_counters.restype = i32
    lib.mpr_debug_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.mpr_last_launch_count.restype = i32
    lib.mpr_last_launch_count.argtypes = [vp]
    lib.mpr_profile_launch_ms.restype = i32
    lib.mpr_profile_launch_ms.argtypes = [vp, i32, C.POINTER(C.c_float)]
    if lib.mpr_abi_version() != ABI_VERSION:
        raise NativeError(f"{LIB_PATH} has ABI version {lib.mpr_abi_version()}, this package needs {ABI_VERSION}: rebuild it "
                          "with `python -m multimodalpromptretrieval_b200.csrc.build --force`")
    _lib = lib
    return lib


class Handle:
    """Owns one ``mpr_handle_t`` (one per process and device)."""

    def __init__(self, device: int):
        self.lib = load()
        self._h = C.c_void_p()
        rc = self.lib.mpr_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = self.lib.mpr_last_error(None)
            raise NativeError(f"mpr_create(device={device}) failed ({rc}): {msg.decode() if msg else '?'}")
        self.device = int(device)

    @property
    def ptr(self) -> C.c_void_p:
        return self._h

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.mpr_last_error(self._h)
            raise NativeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")

    def device_error(self) -> int:
        code = C.c_int(0)
        self.check(self.lib.mpr_device_error(self._h, C.byref(code)), "mpr_device_error")
        return code.value

    def close(self) -> None:
        if self._h:
            self.lib.mpr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
