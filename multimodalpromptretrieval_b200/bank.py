"""``RetrievalBank`` — host mirror of the retrieval half of the reference's ``VQADataset``.

Same method names, arguments, return shapes and attributes as
``create_retrieval_dataset`` (/root/reference/dataset/VQAFeatureDataset.py:118-185) and
``retrieve_closest_qa_pairs`` (:187-246), so ``main.py`` can pass ``bank.retrieve_closest_qa_pairs`` as the model's
``retrieval_function`` (/root/reference/main.py:123, architectures/T5VisionModel.py:143-147) unchanged.  All of the
arithmetic — row cast/normalise, scoring, top-k, merge, vote, token gather — runs in ``libmpr_b200.so``; CLIP stays
stock PyTorch and is reached through the same ``clip_model.encode_image / encode_text`` calls the reference makes.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * the bank is stored bf16 (``retrieval_embeddings`` is this rank's bf16 shard) — scores agree to 1e-3;
  * exact ties resolve to the lower row index (the reference's unstable argsort leaves them undefined);
  * ``use_additional_data`` works (the reference crashes at :181) and extends the info dict per key.
"""
from __future__ import annotations

import hashlib
import json
import os
import pickle
import threading
import weakref
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from . import kernels as K
from .prompt import PromptTables, bucket_lut, prompt_string
from .sharding import CandidateExchange, P2PExchange, plan_shard_reads, shard_bounds

_INFO_KEYS = ("question_type", "question_id", "question")


def _atomic_write(path: str, write) -> None:
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        write(f)
    os.replace(tmp, path)


class LazyPart:
    """A block of ``n_rows`` bank rows produced on demand by ``fn() -> (a, b_or_None)``; only called on the ranks
    whose shard overlaps the block (lets every rank of a sharded job materialise just its own rows)."""

    def __init__(self, n_rows: int, dim: int, fn):
        self.n_rows, self.dim, self.fn = int(n_rows), int(dim), fn


class _AnswerView:
    """``retrieval_answers`` for banks installed from pre-interned ids: a read-only sequence of strings."""

    def __init__(self, ids: np.ndarray, strings: Sequence[str]):
        self.ids, self.strings = ids, strings

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        return self.strings[int(self.ids[i])]


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class _Step:
    """Pre-allocated buffers and the C argument block (``mpr_retrieve_args``) for retrieval steps of ONE shape
    (batch, query halves, dtype, k + skip).  A step is one ctypes call and — for batches that fit one wave of CTAs — one
    kernel launch; everything it needs on the device lives here:

      * staging for host-resident query halves and for the host-tokenised prefix CSR (pinned + device, double-buffered),
      * the scan workspace,
      * TWO device result blocks used alternately (a result stays valid while the next step runs) and their pinned
        host mirrors; a block is laid out ``[status | out_len | majority | count | bucket | q_bias | idx | score | answers |
        keys | input_ids | attention_mask]`` so that ONE device-to-host copy of its head brings a whole step back.
    """

    def __init__(self, bank: "RetrievalBank", b: int, d0: int, d1: int, dtype: torch.dtype, kk: int, skip: int):
        dev = bank.device
        self.bank, self.b, self.d0, self.d1, self.dtype, self.kk, self.skip = bank, b, d0, d1, dtype, kk, skip
        self.k = k = kk - skip
        d = d0 + d1
        # device staging for host-resident query halves, one pair per turn (a queued step reads its own pair)
        self.q0 = [torch.empty((b, d0), dtype=dtype, device=dev) for _ in range(2)]
        self.q1 = [torch.empty((b, d1), dtype=dtype, device=dev) if d1 else None for _ in range(2)]
        # bf16 copy of raw queries for the scan variants that take prepared rows (hybrid q-tile: 512 < D <= 1024 with more
        # than 16 queries; CTA-pair multicast for wider rows)
        self.q_scratch = torch.empty((b, d), dtype=torch.bfloat16, device=dev) if d > 512 and b > 16 else None
        n_local = bank.retrieval_embeddings.shape[0]
        need = K.search_workspace_bytes(b, n_local, d, kk, dev.index)
        if need == 0:
            K.search_plan(b, n_local, d, kk, dev.index)          # raises with the library's precise message
        self.workspace = K.new_workspace(need, dev)
        self.max_len = bank.max_source_length
        # ---- result block layout
        off, lay = 0, {}
        for name, count, size in (("status", 4, 4), ("length", b, 4), ("majority_answer", b, 4), ("majority_count", b, 4),
                                  ("bucket", b, 4), ("q_bias", b, 4), ("idx", b * kk, 4), ("score", b * kk, 4),
                                  ("answer_ids", b * k, 4), ("keys", b * kk, 8)):
            lay[name] = (off, count * size)
            off = _align(off + count * size)
        self.layout, self.head_bytes = lay, off
        self.block_bytes = off + 2 * b * self.max_len * 8
        self.d_blocks = [torch.zeros(self.block_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.h_blocks = [torch.zeros(self.block_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.turn = 0
        # ---- prefix staging (grown on demand), double-buffered with an event each
        self.prefix_cap = 0
        self.h_pre: list = []
        self.d_pre: list = []
        self.pre_done = [None, None]
        self.zero_off = torch.zeros(b + 1, dtype=torch.int32, device=dev)      # "no prefix": vote only
        self.dummy = torch.zeros(16, dtype=torch.int32, device=dev)
        self.h_q = [None, None]                           # pinned staging for pageable host queries, per turn
        self.done_events = [torch.cuda.Event(), torch.cuda.Event()]   # submit(): "result block `turn` is on the host"
        self.in_flight = [False, False]                   # a submitted step still owns this turn's buffers
        self._pinned_cache: Dict[tuple, bool] = {}
        self.stage_busy = [None, None]                    # event after a non-blocking step that read this turn's staging
        self.gen = [0, 0]                                 # submissions per turn (a stale handle must not release a newer step)
        # ---- argument blocks: everything that never changes is filled once
        self.args = _native.RetrieveArgs()
        a = self.args
        a.d0, a.d1, a.q_dtype, a.normalise = d0, d1, K._DTYPES[dtype], int(bank.normalise)
        a.q_scratch = 0 if self.q_scratch is None else self.q_scratch.data_ptr()
        a.b, a.d, a.kk, a.skip = b, d, kk, skip
        a.bank = bank.retrieval_embeddings.data_ptr() if n_local else 0
        a.bias = bank.bias.data_ptr() if n_local else 0
        a.n_local, a.idx_base = n_local, bank.row_begin
        a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        a.answer_id = bank.answer_id.data_ptr()
        a.bucket_lut = bank._lut(k).data_ptr()
        a.max_len = self.max_len
        if bank._p2p is not None and bank._p2p.world_size > 1:
            bank._p2p.fill_args(a)
        self.io = _native.HostIO()
        self._q0_ptr = [t.data_ptr() for t in self.q0]
        self._q1_ptr = [0 if t is None else t.data_ptr() for t in self.q1]
        self._view_cache: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._prompt_mode = -1            # which prompt tables the argument block currently points at
        self._tail_bound: Dict[bool, int] = {}
        self._bound_base = 0
        self.h_pre_np: list = []

    def _is_pinned(self, t: torch.Tensor) -> bool:
        """``Tensor.is_pinned`` asks the driver (~10-30 us); data loaders recycle their pinned buffers, so the answer is
        remembered per (address, size).  A buffer freed and re-allocated pageable at the same address would be misjudged
        only in the harmless direction of an extra lookup miss: entries are dropped when the table is full."""
        key = (t.data_ptr(), t.numel() * t.element_size())
        hit = self._pinned_cache.get(key)
        if hit is None:
            if len(self._pinned_cache) > 64:
                self._pinned_cache.clear()
            hit = bool(t.is_pinned())
            self._pinned_cache[key] = hit
        return hit

    def _views(self, block: torch.Tensor, stride: int) -> Dict[str, torch.Tensor]:
        b, kk, k = self.b, self.kk, self.k
        lay = self.layout

        def v(name, dtype, *shape):
            o, n = lay[name]
            return block[o:o + n].view(dtype).view(*shape)

        big = self.head_bytes
        return {"status": v("status", torch.int32, 4), "length": v("length", torch.int32, b),
                "majority_answer": v("majority_answer", torch.int32, b), "majority_count": v("majority_count", torch.int32, b),
                "bucket": v("bucket", torch.int32, b), "q_bias": v("q_bias", torch.float32, b),
                "idx": v("idx", torch.int32, b, kk), "score": v("score", torch.float32, b, kk),
                "answer_ids": v("answer_ids", torch.int32, b, k), "keys": v("keys", torch.int64, b, kk),
                "input_ids": block[big:big + b * stride * 8].view(torch.int64).view(b, stride),
                "attention_mask": block[big + b * stride * 8:big + 2 * b * stride * 8].view(torch.int64).view(b, stride)}

    def _cached_views(self, turn: int, stride: int, host: bool) -> Dict[str, torch.Tensor]:
        key = (turn, stride, host)
        v = self._view_cache.get(key)
        if v is None:
            if len(self._view_cache) > 256:
                self._view_cache.clear()
            v = self._views((self.h_blocks if host else self.d_blocks)[turn], stride)
            if host:       # numpy mirrors of the words the host looks at on every step
                v["_status_np"] = v["status"].numpy()
                v["_length_np"] = v["length"].numpy()
            self._view_cache[key] = v
        return v

    def _grow_prefix(self, n_ids: int) -> None:
        cap = max(2 * n_ids, 4096)
        dev = self.bank.device
        self.h_pre = [(torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(self.b + 1, dtype=torch.int32).pin_memory())
                      for _ in range(2)]
        self.h_pre_np = [(a.numpy(), o.numpy()) for a, o in self.h_pre]
        self.d_pre = [(torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(self.b + 1, dtype=torch.int32, device=dev))
                      for _ in range(2)]
        self.pre_done = [None, None]
        self.prefix_cap = cap

    def _bind_turn(self, turn: int) -> None:
        """Point the argument block at result block `turn` (everything that does not depend on the row pitch)."""
        a, lay = self.args, self.layout
        base = self.d_blocks[turn].data_ptr()
        a.status = base + lay["status"][0]
        a.out_len = base + lay["length"][0]
        a.maj_answer = base + lay["majority_answer"][0]
        a.maj_count = base + lay["majority_count"][0]
        a.bucket = base + lay["bucket"][0]
        a.out_q_bias = base + lay["q_bias"][0]
        a.out_idx = base + lay["idx"][0]
        a.out_score = base + lay["score"][0]
        a.ret_answer = base + lay["answer_ids"][0]
        a.out_keys = base + lay["keys"][0]
        a.input_ids = base + self.head_bytes
        self._bound_base = base

    def run(self, img, txt, prefix, use_quantifier: bool, to_host: bool, defer: bool = False) -> Dict[str, object]:
        """``img`` / ``txt``: query halves, device tensors or (pinned) host tensors; ``prefix`` = None (vote only) or the
        host CSR ``(ids int32, off int32[b+1])`` of the per-query prefix tokens, or device tensors ``(ids, off, longest)``.
        Returns the result views (device; plus host views when ``to_host``, in which case the call has synchronised).
        ``to_host="async"``: the device-to-host copy is queued but NOT awaited; ``out["done"]`` is the event that follows
        it and ``out["host"]`` must not be read before that event has completed (see :class:`PendingRetrieval`)."""
        bank, a, io = self.bank, self.args, self.io
        dev = bank.device
        turn = self.turn
        self.turn ^= 1
        if self.in_flight[turn]:             # the submitted step that last used this turn's buffers must be through
            self.done_events[turn].synchronize()
            self.in_flight[turn] = False
        self._bind_turn(turn)
        if not img.is_cuda:
            if not self._is_pinned(img):     # pageable host memory: stage through a pinned buffer of our own
                if self.h_q[turn] is None:
                    self.h_q[turn] = (torch.empty((self.b, self.d0), dtype=self.dtype).pin_memory(),
                                      torch.empty((self.b, self.d1), dtype=self.dtype).pin_memory() if self.d1 else None)
                self.h_q[turn][0].copy_(img)
                img = self.h_q[turn][0]
                if txt is not None:
                    self.h_q[turn][1].copy_(txt)
                    txt = self.h_q[turn][1]
            io.h_q0, io.h_q1 = img.data_ptr(), (0 if txt is None else txt.data_ptr())
            a.q0, a.q1 = self._q0_ptr[turn], self._q1_ptr[turn]
        else:
            io.h_q0 = io.h_q1 = 0
            a.q0, a.q1 = img.data_ptr(), (0 if txt is None else txt.data_ptr())
        # ---- prompt stage inputs
        staged_prefix = False
        if prefix is None:
            stride = 1
            if self._prompt_mode != 0:
                a.prefix_ids, a.prefix_off = self.dummy.data_ptr(), self.zero_off.data_ptr()
                a.seg_ids, a.seg_off = self.dummy.data_ptr(), bank._zero_seg_off().data_ptr()
                a.pad_id, a.eos_id = 0, 1
                self._prompt_mode = 0
            io.h_prefix_ids = io.h_prefix_off = 0
            io.n_prefix_ids = 0
        else:
            tables = bank._prompt_tables()
            if self._prompt_mode != 1:
                a.seg_ids, a.seg_off = tables.seg_ids.data_ptr(), tables.seg_off.data_ptr()
                a.pad_id, a.eos_id = tables.pad_id, tables.eos_id
                self._prompt_mode = 1
                self._tail_bound = {True: tables.tail_bound(True), False: tables.tail_bound(False)}
            ids, off = prefix[0], prefix[1]
            if isinstance(ids, torch.Tensor):
                longest = int(prefix[2])
                a.prefix_ids, a.prefix_off = ids.data_ptr(), off.data_ptr()
                io.h_prefix_ids = io.h_prefix_off = 0
                io.n_prefix_ids = 0
            else:
                n_ids = int(ids.size)
                longest = int((off[1:] - off[:-1]).max()) if self.b else 0
                if n_ids > self.prefix_cap:
                    self._grow_prefix(n_ids)
                if self.pre_done[turn] is not None:
                    self.pre_done[turn].synchronize()       # the H2D that last read this pinned pair has finished
                    self.pre_done[turn] = None
                h_ids, h_off = self.h_pre[turn]
                np_ids, np_off = self.h_pre_np[turn]
                d_ids, d_off = self.d_pre[turn]
                np_ids[:n_ids] = ids
                np_off[:] = off
                io.h_prefix_ids, io.h_prefix_off, io.n_prefix_ids = h_ids.data_ptr(), h_off.data_ptr(), n_ids
                a.prefix_ids, a.prefix_off = d_ids.data_ptr(), d_off.data_ptr()
                staged_prefix = True
            stride = min(self.max_len, longest + self._tail_bound[bool(use_quantifier)])
        a.use_quantifier, a.out_stride = int(bool(use_quantifier)), stride
        a.attention_mask = self._bound_base + self.head_bytes + self.b * stride * 8
        if to_host:
            io.d_out, io.h_out = self._bound_base, self.h_blocks[turn].data_ptr()
            io.out_bytes = self.head_bytes + 2 * self.b * stride * 8
            io.sync = 0 if to_host == "async" else 1
        else:
            io.d_out = io.h_out = 0
            io.out_bytes, io.sync = 0, 0
        # sharded bank, on request: leave the collection of the peers' candidates, the vote and the outputs to the finish
        # kernel on the library's side stream (the NVLink latency then hides under the next step's scan); the caller
        # joins (RetrievalBank.join), a submitted step's result copy waits for the finish by itself.  Not the default for
        # submitted steps: deferral costs five more driver calls per step, and at 8 GPUs the host side of a 220 us step
        # is what bounds the two-deep pipeline (measured: 574 k q/s without, 519 k with).
        a.defer_finish = 1 if (defer or (to_host == "async" and bank.defer_submitted_finish)) and a.world > 1 else 0
        copy_streams = bank._copy_streams() if to_host == "async" else None
        if copy_streams is not None:
            # Inputs on st_in, results on st_out.  The input copy overwrites this turn's device staging: the submitted
            # step that last read it was awaited above (in_flight), a blocking call has returned, and a non-blocking
            # call with host inputs left an event behind (stage_busy) that the input stream waits for.
            if self.stage_busy[turn] is not None:
                copy_streams[0].wait_event(self.stage_busy[turn])
                self.stage_busy[turn] = None
            io.stream_in, io.stream_out = copy_streams[0].cuda_stream, copy_streams[1].cuda_stream
        else:
            io.stream_in = io.stream_out = 0
        K.retrieve(a, dev, io)
        if not to_host and (io.h_q0 or staged_prefix):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self.stage_busy[turn] = ev
        out: Dict[str, object] = {"device": self._cached_views(turn, stride, False), "stride": stride}
        if to_host == "async":
            ev = self.done_events[turn]
            ev.record(copy_streams[1])
            self.in_flight[turn] = True
            self.gen[turn] += 1
            out["owner"] = (self, turn, self.gen[turn])
            if staged_prefix:
                self.pre_done[turn] = ev
            out["done"] = ev
            out["host"] = self._cached_views(turn, stride, True)
            return out
        if staged_prefix and not to_host:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self.pre_done[turn] = ev
        if to_host:
            hv = self._cached_views(turn, stride, True)
            status = int(hv["_status_np"][0])
            if status != 0:
                raise K._native.NativeError(f"retrieval step reported device status {status} "
                                            f"({'a peer rank did not deliver its candidates in time' if status == _native.STATUS_XCHG_TIMEOUT else 'see mpr_b200.h'})")
            out["host"] = hv
        return out


class PendingRetrieval:
    """A retrieval step that has been queued on the GPU (:meth:`RetrievalBank.submit_prompt_ids_host`): inputs copied in,
    kernel launched, result copy queued.  :meth:`result` waits for exactly this step and returns what
    :meth:`RetrievalBank.retrieve_prompt_ids_host` returns.  At most TWO steps of a shape may be outstanding: the result
    buffers alternate, so the step after next reuses this one's."""

    __slots__ = ("_step", "_done")

    def __init__(self, step: Dict[str, object]):
        self._step = step
        self._done = False

    def wait(self) -> Dict[str, torch.Tensor]:
        """Host views of the whole result block (idx, score, vote, input_ids, ...), valid until the step after next."""
        host = self._step["host"]
        if not self._done:
            self._done = True
            ev = self._step.get("done")
            if ev is None:             # the NCCL-exchange path has already synchronised
                return host
            ev.synchronize()
            owner = self._step.get("owner")
            if owner is not None and owner[0].gen[owner[1]] == owner[2]:
                owner[0].in_flight[owner[1]] = False
            status = int(host["_status_np"][0])
            if status != 0:
                raise K._native.NativeError(
                    f"retrieval step reported device status {status} "
                    f"({'a peer rank did not deliver its candidates in time' if status == _native.STATUS_XCHG_TIMEOUT else 'see mpr_b200.h'})")
        return host

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        host = self.wait()
        length = host.get("_length_np")
        longest_out = int(length.max()) if length is not None else int(host["length"].max())
        return host["input_ids"][:, :longest_out], host["attention_mask"][:, :longest_out]


class RetrievalBank:
    def __init__(self, clip_model=None, clip_tokenize=None, tokenizer=None, device=None, normalise: bool = False,
                 process_group=None, shard: bool = True, max_source_length: int = 512, name: str = "VQADataset",
                 cache_root: str = "cache", additional_root: str = os.path.join("synthetic_data", "cache",
                                                                               "ROCOFeatureDataset"),
                 memoise: bool = True, use_cuda_graph: bool = False, exchange: str = "p2p",
                 fuse_query_cast: bool = True, precomputed_features: bool = False, exchange_capacity: int = 65536):
        if not torch.cuda.is_available():
            raise RuntimeError("RetrievalBank needs a B200 (sm_100a) GPU; there is no CPU fallback path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.clip_model = clip_model
        self.clip_tokenize = clip_tokenize
        self.tokenizer = tokenizer
        self.normalise = bool(normalise)
        self.max_source_length = int(max_source_length)
        self.name = name
        self.cache_root = cache_root
        self.additional_root = additional_root
        self.memoise = memoise
        # use_cuda_graph / fuse_query_cast are accepted for compatibility: a step is a single launch with the query
        # preparation inside it, so there is no chain left to capture or to fuse
        self.use_cuda_graph = bool(use_cuda_graph)
        self.fuse_query_cast = bool(fuse_query_cast)
        self.precomputed_features = bool(precomputed_features)
        self.exchange_capacity = int(exchange_capacity)
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' (all-gather + merge) or 'p2p' (peer-memory push of epoch-tagged words inside the retrieval kernel)")
        self.exchange_mode = exchange
        self._p2p: Optional[P2PExchange] = None
        self._process_group = process_group if shard else None
        self._steps: Dict[tuple, _Step] = {}
        self._zero_seg: Optional[torch.Tensor] = None
        self._pool = None
        self._copy_stream_pair = None
        self.defer_submitted_finish = False      # see _Step.run
        self._prefetched: Dict[tuple, object] = {}
        self._prefetch_lock = threading.Lock()
        self.exchange = CandidateExchange(process_group if shard else None)
        if not shard:
            self.exchange.rank, self.exchange.world_size = 0, 1
        # state the reference keeps on the dataset object (VQAFeatureDataset.py:119-120,128-134,159-161)
        self.is_training_phase = True
        self.retrieval_k = 15
        self.retrieval_embeddings: Optional[torch.Tensor] = None     # bf16 [n_local, D] (this rank's rows)
        self.retrieval_answers: List[str] = []
        self.retrieval_question_info: Dict[str, List[str]] = {}
        # device-side tables
        self.bias: Optional[torch.Tensor] = None                      # fp32 [n_local] = -0.5*|row|^2
        self.answer_id: Optional[torch.Tensor] = None                 # int32 [N] interned answers, replicated
        self.answer_strings: List[str] = []
        self.n_total = 0
        self.row_begin = 0
        self._tables: Optional[PromptTables] = None
        self._lut_cache: Dict[int, torch.Tensor] = {}
        self._memo = None
        K.handle(self.device.index)   # fail loudly now if the library / device is unusable

    # ------------------------------------------------------------------------------------------ bank build
    def create_retrieval_dataset(self, data_loader, prefix, is_training_phase: bool = True, retrieval_k: int = 15,
                                 use_additional_data: bool = False) -> None:
        """Same contract as VQAFeatureDataset.py:118-185 (``prefix`` is accepted and unused, as in the reference)."""
        self._check_kk(retrieval_k + (1 if is_training_phase else 0))
        self.is_training_phase = is_training_phase
        self.retrieval_k = retrieval_k
        cache_dir = os.path.join(self.cache_root, self.name)
        embedding_path = os.path.join(cache_dir, "embedding.pt")
        question_info_path = os.path.join(cache_dir, "answer_types.pkl")
        answer_path = os.path.join(cache_dir, "answers.pkl")

        parts: List[Tuple[torch.Tensor, Optional[torch.Tensor]]] = []
        if os.path.exists(embedding_path) and os.path.exists(answer_path):
            emb = torch.load(embedding_path, map_location="cpu")
            print(f"Loaded cached qa lookup embeddings from {embedding_path} ...")
            with open(answer_path, "rb") as f:
                answers = pickle.load(f)
                print(f"Loaded cached qa lookup answers from {answer_path} ...")
            with open(question_info_path, "rb") as f:
                info = pickle.load(f)
                print(f"Loaded cached qa lookup answer types from {question_info_path} ...")
            parts.append((emb, None))
        else:
            os.makedirs(cache_dir, exist_ok=True)
            print(f"Creating qa pairs in {cache_dir} ...")
            answers, info = [], {k: [] for k in _INFO_KEYS}
            host_rows = []
            with torch.no_grad():
                for batch in data_loader:
                    img, txt = self._encode(batch)
                    parts.append((img, txt))
                    host_rows.append(torch.cat([img, txt], 1).float().cpu())
                    answers.extend(batch["answer"])
                    info["question_type"].extend(batch["question_type"])
                    info["question_id"].extend(batch["question_id"])
                    info["question"].extend(batch["question"])
            # Every rank walked its own data_loader: the row order (and with it the shard boundaries and the replicated
            # answer table) must be IDENTICAL everywhere — the reference's retrieval loader shuffles (main.py:105), so
            # the ranks' sampler seeds must agree.
            self._check_rank_consistency(info["question_id"], answers)
            if self.exchange.rank == 0:   # reference-format cache (fp32 [N, 1024] + two pickles), :163-167
                # the embedding file is what the other ranks' os.path.exists test looks for: it is written last, and
                # every file appears atomically (temp file + rename)
                _atomic_write(answer_path, lambda f: pickle.dump(answers, f))
                _atomic_write(question_info_path, lambda f: pickle.dump(info, f))
                _atomic_write(embedding_path, lambda f: torch.save(torch.cat(host_rows, 0) if host_rows else torch.zeros(0, 0), f))
            self._barrier()

        if use_additional_data:   # :169-181, with the dict.extend bug at :181 replaced by a per-key extend
            roco_feats = torch.load(os.path.join(self.additional_root, "embedding.pt"), map_location="cpu")
            with open(os.path.join(self.additional_root, "answers.pkl"), "rb") as f:
                roco_ans = pickle.load(f)
            with open(os.path.join(self.additional_root, "answer_types.pkl"), "rb") as f:
                roco_info = pickle.load(f)
            parts.append((roco_feats, None))
            answers = list(answers) + list(roco_ans)
            info = {k: list(v) + list(roco_info.get(k, [])) for k, v in info.items()}

        self.install_bank(parts, answers, info)
        print(f"Retrieval features shape: {torch.Size([self.n_total, self.dim])}")
        print(f"Number of answers: {len(self.retrieval_answers)}")

    def _barrier(self) -> None:
        import torch.distributed as dist
        if self.exchange.world_size > 1 and dist.is_available() and dist.is_initialized():
            dist.barrier(self._process_group)

    def _check_rank_consistency(self, *sequences) -> None:
        """Raises on every rank if the ranks disagree about the bank's row order / contents (hash of the sequences)."""
        import torch.distributed as dist
        if self.exchange.world_size <= 1 or not (dist.is_available() and dist.is_initialized()):
            return
        hsh = hashlib.sha1()
        for seq in sequences:
            hsh.update(str(len(seq)).encode())
            for x in seq:
                hsh.update(str(x).encode())
                hsh.update(b"\0")
        mine = hsh.hexdigest()
        everyone = [None] * self.exchange.world_size
        dist.all_gather_object(everyone, mine, group=self._process_group)
        if len(set(everyone)) != 1:
            raise RuntimeError("the ranks built DIFFERENT retrieval banks (row order or contents differ: "
                               f"{everyone}); give every rank's retrieval DataLoader the same sampler seed, or build on "
                               "one rank and load the cache / shard directory on the others")

    def install_bank(self, parts: Iterable, answers: Optional[Sequence[str]], info: Optional[Dict[str, Sequence[str]]],
                     is_training_phase: Optional[bool] = None, retrieval_k: Optional[int] = None,
                     answer_ids: Optional[np.ndarray] = None, answer_strings: Optional[Sequence[str]] = None) -> None:
        """Lays the bank out in HBM.  ``parts`` is a sequence of row blocks in global row order, each either
        ``(combined [n, D], None)``, ``(image_half [n, d0], text_half [n, d1])`` (fp32/fp16/bf16, any device) or a
        :class:`LazyPart`.  Kernel 1 casts (and optionally normalises) every block straight into this rank's bf16
        shard.  Answers are either a list of N strings or pre-interned (``answer_ids`` int32 [N] + ``answer_strings``)."""
        if is_training_phase is not None:
            self.is_training_phase = is_training_phase
        if retrieval_k is not None:
            self.retrieval_k = retrieval_k
        with K.nvtx_range("mpr.bank_build"):
            self._install_bank(parts, answers, info, answer_ids, answer_strings)

    def _install_bank(self, parts, answers, info, answer_ids, answer_strings) -> None:
        parts = list(parts)
        rows_of = lambda p: p.n_rows if isinstance(p, LazyPart) else int(p[0].shape[0])
        n_total = sum(rows_of(p) for p in parts)
        if n_total == 0:
            raise ValueError("empty retrieval bank")
        p0 = parts[0]
        dim = p0.dim if isinstance(p0, LazyPart) else int(p0[0].shape[1] + (p0[1].shape[1] if p0[1] is not None else 0))
        n_answers = len(answers) if answers is not None else len(answer_ids)
        if n_answers != n_total:
            raise ValueError(f"{n_answers} answers for {n_total} bank rows")
        begin, end = shard_bounds(n_total, self.exchange.rank, self.exchange.world_size)
        n_local = end - begin
        bank = torch.empty((max(n_local, 1), dim), dtype=torch.bfloat16, device=self.device)[:n_local]
        bias = torch.empty((max(n_local, 1),), dtype=torch.float32, device=self.device)[:n_local]
        row = 0
        chunk_rows = max(1, (256 << 20) // (dim * 4))       # stage host blocks through <= 256 MiB device buffers
        for part in parts:
            n = rows_of(part)
            lo, hi = max(begin, row), min(end, row + n)
            if lo < hi:
                a, b = part.fn() if isinstance(part, LazyPart) else part
                for c0 in range(lo, hi, chunk_rows):
                    c1 = min(hi, c0 + chunk_rows)
                    sa = a[c0 - row:c1 - row].to(self.device, non_blocking=True).contiguous()
                    sb = None if b is None else b[c0 - row:c1 - row].to(self.device, non_blocking=True).contiguous()
                    if sa.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                        sa = sa.float()
                        sb = None if sb is None else sb.float()
                    K.bank_build(sa, sb, normalise=self.normalise, out=bank[c0 - begin:c1 - begin],
                                 bias=bias[c0 - begin:c1 - begin])
                del a, b
            row += n
        self.retrieval_embeddings = bank
        self.bias = bias
        self.n_total, self.dim, self.row_begin = n_total, dim, begin
        self.retrieval_question_info = {k: list(v) for k, v in (info or {}).items()}
        if answers is not None:
            # intern the answers: bank row -> answer id, replicated on every rank (4 B/row)
            self.retrieval_answers = list(answers)
            table: Dict[str, int] = {}
            ids = np.empty(n_total, dtype=np.int32)
            for i, a in enumerate(self.retrieval_answers):
                ids[i] = table.setdefault(a, len(table))
            self.answer_strings = list(table.keys())
        else:
            ids = np.ascontiguousarray(answer_ids, dtype=np.int32)
            self.answer_strings = list(answer_strings)
            self.retrieval_answers = _AnswerView(ids, self.answer_strings)
        self.answer_id = torch.from_numpy(ids).to(self.device)
        self._tables = None
        self._memo = None
        self._steps = {}

    # ------------------------------------------------------------------------------------------ shard cache (N2)
    SHARD_FORMAT = 1

    @staticmethod
    def cache_key(**parts) -> str:
        """Key for the shard cache.  The reference keys its cache on the dataset CLASS NAME only
        (VQAFeatureDataset.py:122-124), so a changed split / subset / union / CLIP checkpoint silently reuses a stale
        bank; here the caller lists what the bank depends on and the manifest refuses a mismatch."""
        return hashlib.sha1(json.dumps(parts, sort_keys=True, default=str).encode()).hexdigest()[:16]

    def save_shards(self, directory: str, key: str = "") -> None:
        """Writes this rank's rows exactly as they sit in HBM — ``shard_<rank>.bf16`` (raw bf16 row-major) and
        ``bias_<rank>.f32`` — plus, on rank 0, the interned answers, the info dict and ``manifest.json``.  Loading
        needs no cast and no CLIP pass, and works for any later GPU count (files are addressed by row range)."""
        os.makedirs(directory, exist_ok=True)
        r, w = self.exchange.rank, self.exchange.world_size
        self.retrieval_embeddings.view(torch.int16).cpu().numpy().tofile(os.path.join(directory, f"shard_{r:03d}.bf16"))
        self.bias.cpu().numpy().tofile(os.path.join(directory, f"bias_{r:03d}.f32"))
        self._barrier()                    # every rank's shard file is complete before the manifest names it
        if r == 0:
            self.answer_id.cpu().numpy().tofile(os.path.join(directory, "answer_id.i32"))
            _atomic_write(os.path.join(directory, "meta.pkl"), lambda f: pickle.dump(
                {"answer_strings": self.answer_strings, "info": self.retrieval_question_info}, f))
            manifest = {"format": self.SHARD_FORMAT, "key": key, "n_total": self.n_total, "dim": self.dim,
                        "normalise": self.normalise, "world_size": w,
                        "ranges": [list(shard_bounds(self.n_total, i, w)) for i in range(w)]}
            _atomic_write(os.path.join(directory, "manifest.json"), lambda f: f.write(json.dumps(manifest).encode()))
        self._barrier()                    # the manifest exists when save_shards returns, on every rank

    def load_shards(self, directory: str, key: str = "", is_training_phase: Optional[bool] = None,
                    retrieval_k: Optional[int] = None) -> bool:
        """Loads this rank's rows from a shard directory written by :meth:`save_shards` under ANY world size.
        Returns False (and loads nothing) if there is no manifest or it does not match ``key`` / ``normalise``."""
        path = os.path.join(directory, "manifest.json")
        if not os.path.exists(path):
            return False
        with open(path) as f:
            m = json.load(f)
        if m.get("format") != self.SHARD_FORMAT or m.get("key", "") != key or bool(m["normalise"]) != self.normalise:
            return False
        n_total, dim = int(m["n_total"]), int(m["dim"])
        begin, end = shard_bounds(n_total, self.exchange.rank, self.exchange.world_size)
        n_local = end - begin
        bank = torch.empty((max(n_local, 1), dim), dtype=torch.bfloat16, device=self.device)[:n_local]
        bias = torch.empty((max(n_local, 1),), dtype=torch.float32, device=self.device)[:n_local]
        for i, first, n, dst in plan_shard_reads([tuple(x) for x in m["ranges"]], begin, end):
            rows = np.memmap(os.path.join(directory, f"shard_{i:03d}.bf16"), dtype=np.int16, mode="r",
                             offset=first * dim * 2, shape=(n, dim))
            bank[dst:dst + n].view(torch.int16).copy_(torch.from_numpy(np.ascontiguousarray(rows)))
            b = np.memmap(os.path.join(directory, f"bias_{i:03d}.f32"), dtype=np.float32, mode="r",
                          offset=first * 4, shape=(n,))
            bias[dst:dst + n].copy_(torch.from_numpy(np.ascontiguousarray(b)))
        with open(os.path.join(directory, "meta.pkl"), "rb") as f:
            meta = pickle.load(f)
        ids = np.fromfile(os.path.join(directory, "answer_id.i32"), dtype=np.int32)
        if is_training_phase is not None:
            self.is_training_phase = is_training_phase
        if retrieval_k is not None:
            self.retrieval_k = retrieval_k
        self.retrieval_embeddings, self.bias = bank, bias
        self.n_total, self.dim, self.row_begin = n_total, dim, begin
        self.answer_strings = list(meta["answer_strings"])
        self.retrieval_question_info = {k: list(v) for k, v in meta["info"].items()}
        self.retrieval_answers = _AnswerView(ids, self.answer_strings)
        self.answer_id = torch.from_numpy(ids).to(self.device)
        self._tables, self._memo, self._steps = None, None, {}
        return True

    # ------------------------------------------------------------------------------------------ query path
    def _encode(self, batch) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """The two CLIP calls of VQAFeatureDataset.py:146-147 / :189-190 (stock PyTorch, out of scope).  With
        ``precomputed_features=True`` the batch already carries the embeddings (``batch["image"]`` = image half or the
        whole row, optional ``batch["text_embedding"]``) and they are handed to the kernels where they lie — on the
        host they are copied by the retrieval call itself."""
        if self.precomputed_features:
            txt = batch.get("text_embedding") if isinstance(batch, dict) else None
            return batch["image"], txt
        img = self.clip_model.encode_image(batch["image"].to(self.device, non_blocking=True))
        tokens = self.clip_tokenize(batch["question"]) if self.clip_tokenize is not None else batch["question"]
        if isinstance(tokens, torch.Tensor):
            tokens = tokens.to(self.device, non_blocking=True)
        txt = self.clip_model.encode_text(tokens)
        return img.detach().contiguous(), (None if txt is None else txt.detach().contiguous())

    def _copy_streams(self):
        """Input / result copy streams of the two-deep host pipeline (submit_prompt_ids_host): copies of the neighbouring
        steps overlap the current step's kernel."""
        if self._copy_stream_pair is None:
            with torch.cuda.device(self.device):
                self._copy_stream_pair = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return self._copy_stream_pair

    def _zero_seg_off(self) -> torch.Tensor:
        if self._zero_seg is None or self._zero_seg.numel() != 9 + len(self.answer_strings):
            self._zero_seg = torch.zeros(9 + len(self.answer_strings), dtype=torch.int32, device=self.device)
        return self._zero_seg

    def _ensure_exchange(self, need_keys: int) -> None:
        """Sharded banks exchange candidates through peer memory inside the retrieval kernel (exchange="p2p", default)
        or through an NCCL all-gather between two launches (exchange="nccl").  The peer buffer is sized ONCE
        (``exchange_capacity`` keys, default 64 Ki = e.g. 2048 queries x 32): creating it is a collective."""
        if self.exchange.world_size <= 1 or self.exchange_mode != "p2p" or self._p2p is not None:
            return
        cap = max(self.exchange_capacity, need_keys)
        try:
            self._p2p = P2PExchange(self.device, cap, self._process_group)
        except Exception as e:       # no symmetric memory on this box: keep the collective-library path
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({type(e).__name__}: {e}); using the NCCL all-gather path")
            self.exchange_mode = "nccl"

    def _step(self, img: torch.Tensor, txt: Optional[torch.Tensor], kk: int, skip: int) -> _Step:
        if img.dtype not in K._DTYPES:
            raise TypeError(f"query embeddings must be float32/float16/bfloat16, got {img.dtype}")
        if txt is not None and txt.dtype != img.dtype:
            raise TypeError("image and text halves must share a dtype")
        b, d0 = int(img.shape[0]), int(img.shape[1])
        d1 = 0 if txt is None else int(txt.shape[1])
        if d0 + d1 != self.dim:
            raise ValueError(f"query width {d0}+{d1} does not match the bank's {self.dim}")
        self._ensure_exchange(b * kk)
        if self._p2p is not None and b * kk > self._p2p.cap:
            raise ValueError(f"batch x (k+skip) = {b * kk} exceeds exchange_capacity = {self._p2p.cap}; construct the bank "
                             "with a larger exchange_capacity (growing it mid-run would be a hidden collective)")
        key = (b, d0, d1, img.dtype, kk, skip)
        st = self._steps.get(key)
        if st is None:
            st = _Step(self, b, d0, d1, img.dtype, kk, skip)
            self._steps[key] = st
        return st

    def _check_kk(self, kk: int) -> None:
        if not 1 <= kk <= _native.MPR_MAX_KK:
            raise ValueError(f"k + skip must be in [1, {_native.MPR_MAX_KK}] (got {kk}): the per-query lists of the scan "
                             "kernel hold at most 32 candidates; the reference accepts any k")

    def run_step(self, img: torch.Tensor, txt: Optional[torch.Tensor] = None, prefix=None, use_quantifier: bool = True,
                 to_host: bool = False, kk: Optional[int] = None, skip: Optional[int] = None, defer: bool = False
                 ) -> Dict[str, object]:
        """One retrieval step on prepared inputs: query halves (device or host tensors) [+ prefix token CSR] ->
        top-(k+skip), vote, bucket [, prompt ids].  Returns {"device": views, "host": views (to_host), "stride"}; the
        views alias buffers that the step after next overwrites."""
        if skip is None:
            skip = 1 if self.is_training_phase else 0
        if kk is None:
            kk = self.retrieval_k + skip
        self._check_kk(kk)
        img = img.contiguous()
        txt = None if txt is None else txt.contiguous()
        st = self._step(img, txt, kk, skip)
        if self.exchange.world_size > 1 and self.exchange_mode == "nccl":
            return self._run_step_nccl(st, img, txt, prefix, use_quantifier, to_host)
        if torch.cuda.current_device() == self.device.index and not K._NVTX:
            return st.run(img, txt, prefix, use_quantifier, to_host, defer)       # hot path: no context managers
        with torch.cuda.device(self.device), K.nvtx_range("mpr.retrieval_step"):
            return st.run(img, txt, prefix, use_quantifier, to_host, defer)

    def join(self) -> None:
        """After ``run_step(..., defer=True)`` on a sharded bank: the current stream waits for the deferred part of every
        step queued so far (collection of the peers' candidates, merge, vote, prompt ids), after which the device views
        may be consumed on it."""
        with torch.cuda.device(self.device):
            K.retrieve_join(self.device)

    def _run_step_nccl(self, st: _Step, img, txt, prefix, use_quantifier, to_host) -> Dict[str, object]:
        """Sharded step with a collective library between two launches: local scan -> all-gather -> merge -> prompt."""
        img_d = img.to(self.device, non_blocking=True)
        txt_d = None if txt is None else txt.to(self.device, non_blocking=True)
        kk = st.kk
        with torch.cuda.device(self.device):
            keys, _, _, qbias = K.search_topk_fused(img_d, txt_d, self.retrieval_embeddings, self.bias, kk,
                                                    normalise=self.normalise, idx_base=self.row_begin,
                                                    workspace=st.workspace) if self.retrieval_embeddings.shape[0] else \
                (torch.zeros((st.b, kk), dtype=torch.int64, device=self.device), None, None,
                 K.bank_build(img_d, txt_d, normalise=self.normalise)[1])
            keys, score, idx = K.merge_topk(self.exchange.gather(keys))
            tables = self._prompt_tables() if prefix is not None else None
            if prefix is None:
                pre_ids, pre_off, stride = st.dummy, st.zero_off, 1
                seg_ids, seg_off, pad, eos = st.dummy, self._zero_seg_off(), 0, 1
            else:
                ids, off = prefix[0], prefix[1]
                if not (isinstance(ids, torch.Tensor) and ids.is_cuda):
                    longest = int(np.diff(off).max()) if st.b else 0
                    ids = torch.from_numpy(np.ascontiguousarray(ids if ids.size else np.zeros(1, np.int32))).to(self.device)
                    off = torch.from_numpy(np.ascontiguousarray(off)).to(self.device)
                else:
                    longest = int(prefix[2])
                pre_ids, pre_off = ids, off
                stride = min(st.max_len, longest + tables.tail_bound(use_quantifier))
                seg_ids, seg_off, pad, eos = tables.seg_ids, tables.seg_off, tables.pad_id, tables.eos_id
            out = K.prompt_gather(idx, st.skip, self.answer_id, self._lut(st.k), pre_ids, pre_off, seg_ids, seg_off,
                                  use_quantifier, pad, eos, st.max_len, stride)
        out.update({"keys": keys, "score": score, "idx": idx, "q_bias": qbias,
                    "status": torch.zeros(4, dtype=torch.int32, device=self.device)})
        res: Dict[str, object] = {"device": out, "stride": stride}
        if to_host:
            res["host"] = {k_: v.cpu() for k_, v in out.items()}
        return res

    def search_embeddings(self, image_half: torch.Tensor, text_half: Optional[torch.Tensor] = None, kk: Optional[int] = None
                          ) -> Dict[str, torch.Tensor]:
        """Query halves -> global top-(k+skip): ``keys`` u64-as-int64 / ``score`` fp32 / ``idx`` int32 ``[B, kk]`` and
        ``q_bias`` fp32 ``[B]`` (= -0.5*|q|^2), device tensors.  One launch for batches up to one wave of CTAs: query cast,
        scan, merge and — on sharded banks — the peer-memory exchange all run inside it."""
        skip = 1 if self.is_training_phase else 0
        if kk is None:
            kk = self.retrieval_k + skip
        res = self.run_step(image_half, text_half, None, True, False, kk=kk, skip=min(skip, kk - 1))["device"]
        return {"keys": res["keys"], "score": res["score"], "idx": res["idx"], "q_bias": res["q_bias"]}      # |q|^2 = -2 * q_bias

    def _lut(self, k: int) -> torch.Tensor:
        t = self._lut_cache.get(k)
        if t is None:
            t = torch.from_numpy(bucket_lut(k)).to(self.device)
            self._lut_cache[k] = t
        return t

    def _prompt_tables(self) -> PromptTables:
        if self._tables is None:
            if self.tokenizer is None:
                raise RuntimeError("a T5 tokenizer is required for token-id output")
            self._tables = PromptTables(self.tokenizer, self.answer_strings, self.device)
        return self._tables

    # --- host tokenisation of the per-query prefixes, optionally one batch ahead on a worker thread
    def _prefix_csr(self, batch, use_quantifier: bool):
        key = (id(batch["question"]), len(batch["question"]), bool(use_quantifier))
        with self._prefetch_lock:
            fut = self._prefetched.pop(key, None)
        if fut is not None:
            return fut.result()
        with K.nvtx_range("mpr.tokenise"):
            return self._prompt_tables().prefix_tokens(batch["task"], batch["question"], use_quantifier)

    def prefetch(self, batch, use_quantifier: bool = True) -> None:
        """Tokenise the NEXT batch's prefixes on a worker thread while the current step runs on the GPU (they do not
        depend on retrieval).  The data loader hands batches over ahead of time (main.py:176-179); call this with the
        batch object that will later be passed to :meth:`retrieve_prompt_ids` / :meth:`retrieve_prompt_ids_host`."""
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mpr-tokenise")
        tables = self._prompt_tables()
        key = (id(batch["question"]), len(batch["question"]), bool(use_quantifier))
        def job(tasks=batch["task"], questions=batch["question"]):
            with K.nvtx_range("mpr.tokenise_prefetch"):
                return tables.prefix_tokens(tasks, questions, use_quantifier)

        fut = self._pool.submit(job)
        with self._prefetch_lock:
            if len(self._prefetched) > 8:
                self._prefetched.clear()
            self._prefetched[key] = fut

    def _memo_key(self, batch):
        img_t = batch["image"]
        return (self.is_training_phase, self.retrieval_k, getattr(img_t, "_version", 0), tuple(batch["question"]))

    def _retrieve(self, batch, use_quantifier: bool = True, want_ids: bool = False, to_host: bool = False) -> dict:
        """One search per batch object; the reference re-embeds and re-searches the same batch up to 5 times
        (main.py:178-179, 263-270) — results are memoised on the identity and version of ``batch["image"]``, the
        questions and the bank's (is_training_phase, retrieval_k).  Prompt ids asked for later on a memoised search
        only cost the stand-alone token-gather kernel."""
        img_t = batch["image"]
        quant = bool(use_quantifier)
        if self.memoise and self._memo is not None:
            ref, key, out = self._memo
            if ref() is img_t and key == self._memo_key(batch):
                if want_ids and quant not in out["ids"]:
                    tables = self._prompt_tables()
                    pre_ids, pre_off, longest = tables.prefixes(batch["task"], batch["question"], quant)
                    stride = min(self.max_source_length, longest + tables.tail_bound(quant))
                    with torch.cuda.device(self.device):
                        out["ids"][quant] = K.prompt_gather(out["device"]["idx"], out["skip"], self.answer_id,
                                                            self._lut(out["k"]), pre_ids, pre_off, tables.seg_ids,
                                                            tables.seg_off, quant, tables.pad_id, tables.eos_id,
                                                            self.max_source_length, stride)
                return out
        skip = 1 if self.is_training_phase else 0
        k = self.retrieval_k
        self._check_kk(k + skip)
        with torch.no_grad():
            img, txt = self._encode(batch)
        prefix = self._prefix_csr(batch, quant) if want_ids else None
        step = self.run_step(img, txt, prefix, quant, to_host, kk=k + skip, skip=skip)
        # everything stays on the device until somebody asks
        out = {"skip": skip, "k": k, "step": step, "device": step["device"], "ids": {}, "ids_host": {}}
        if want_ids:
            out["ids"][quant] = step["device"]
            if to_host:
                out["ids_host"][quant] = step["host"]
        if self.memoise:
            try:
                self._memo = (weakref.ref(img_t), self._memo_key(batch), out)
            except TypeError:
                self._memo = None
        return out

    @staticmethod
    def _host(r: dict) -> dict:
        """Single D2H of the search result (replaces the reference's B*k ``tensor.__index__`` syncs at :199)."""
        if "idx" not in r:
            src = r["step"].get("host") or r["device"]
            as_np = lambda t: t.cpu().numpy().copy()
            r["idx"] = as_np(src["idx"])
            r["score"] = as_np(src["score"])
            r["q_sqnorm"] = as_np(src["q_bias"]) * -2.0
            r["vote"] = {k_: as_np(src[k_]) for k_ in ("majority_answer", "majority_count", "bucket", "answer_ids")}
        return r

    def retrieve_closest_qa_pairs(self, batch, return_ans: bool = False, return_info=None, return_dists: bool = False,
                                  use_quantifier: bool = True):
        """Same contract as VQAFeatureDataset.py:187-246 (precedence return_ans > return_info > return_dists)."""
        r = self._host(self._retrieve(batch))
        skip, k = r["skip"], r["k"]
        top = r["idx"][:, skip:skip + k]
        answers = [[self.retrieval_answers[int(x)] for x in row if x >= 0] for row in top]            # :199
        if return_ans:
            return answers
        if return_info:                                                                               # :202-210
            out = []
            for row in top:
                info: List[str] = []
                for idx in row:
                    if idx >= 0:
                        info.extend(self.retrieval_question_info[entry][int(idx)] for entry in return_info)
                out.append(info)
            return out
        if return_dists:
            # :243 sorts the whole matrix again and takes ranks 0..k-1 WITHOUT the training skip; reproduced here.
            d2 = r["q_sqnorm"][:, None] - 2.0 * r["score"][:, 0:k]
            dists = np.sqrt(np.maximum(d2, 0.0)).astype(np.float32)
            return list(zip(answers, dists))
        vote = r["vote"]                       # the vote ran in the retrieval kernel's tail
        maj, bkt = vote["majority_answer"], vote["bucket"]
        return [prompt_string(int(bkt[i]), self.answer_strings[int(maj[i])], use_quantifier) for i in range(len(maj))]

    def retrieve_prompt_ids(self, batch, use_quantifier: bool = True, pad_to: str = "longest", copy: bool = True):
        """Additive fast path for ``prepare_input`` (architectures/T5VisionModel.py:143-167): returns the
        ``input_ids`` / ``attention_mask`` the reference's tokenizer call would produce for
        ``task_prefix + question + retrieved_info`` — assembled on the device in the retrieval kernel's tail, no strings
        involved.  Device tensors; ``copy=False`` returns views that the step after next overwrites."""
        r = self._retrieve(batch, use_quantifier, want_ids=True)
        out = r["ids"][bool(use_quantifier)]
        ids, mask = out["input_ids"], out["attention_mask"]
        if pad_to == "longest":        # exact shape parity with padding="longest" costs one tiny D2H sync
            longest_out = int(out["length"].max().item())
            ids, mask = ids[:, :longest_out], mask[:, :longest_out]
        return (ids.clone(), mask.clone()) if copy else (ids, mask)

    def retrieve_prompt_ids_host(self, batch, use_quantifier: bool = True):
        """The same for a HOST consumer, end to end in one library call: host-resident inputs are copied in, the step
        runs, and one device-to-host copy brings ``input_ids`` / ``attention_mask`` (padding="longest" applied) back.
        Returns CPU tensors that alias a pinned buffer reused by the step after next."""
        quant = bool(use_quantifier)
        r = self._retrieve(batch, quant, want_ids=True, to_host=True)
        host = r["ids_host"].get(quant)
        if host is None:               # memoised search whose ids were produced on the device only
            dev = r["ids"][quant]
            host = {k_: dev[k_].cpu() for k_ in ("input_ids", "attention_mask", "length")}
            r["ids_host"][quant] = host
        length = host.get("_length_np")
        longest_out = int(length.max()) if length is not None else int(host["length"].max())
        return host["input_ids"][:, :longest_out], host["attention_mask"][:, :longest_out]

    def submit_prompt_ids_host(self, batch, use_quantifier: bool = True) -> PendingRetrieval:
        """Two-deep software pipeline for a host consumer: queues this batch's whole step (host-to-device copy of the
        embeddings and prefix tokens, the retrieval kernel, the device-to-host copy of the result block) and returns at
        once; ``.result()`` of the returned handle waits for it.  A training loop that gets batch i+1 from its data loader
        while batch i is in flight (/root/reference/main.py:176-179) calls ``nxt = submit(batch_{i+1}); ids, mask =
        cur.result(); cur = nxt`` and the host part of a step hides under the previous step's kernel.  Not memoised; on
        a sharded bank it is a collective like every search (all ranks submit the same batches in the same order)."""
        quant = bool(use_quantifier)
        skip = 1 if self.is_training_phase else 0
        k = self.retrieval_k
        self._check_kk(k + skip)
        with torch.no_grad():
            img, txt = self._encode(batch)
        prefix = self._prefix_csr(batch, quant)
        step = self.run_step(img, txt, prefix, quant, "async", kk=k + skip, skip=skip)
        return PendingRetrieval(step)
