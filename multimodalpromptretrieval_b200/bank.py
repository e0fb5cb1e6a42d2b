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
import weakref
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import kernels as K
from .prompt import PromptTables, bucket_lut, prompt_string
from .sharding import CandidateExchange, P2PExchange, plan_shard_reads, shard_bounds

_INFO_KEYS = ("question_type", "question_id", "question")


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


class _SearchGraph:
    """One captured search chain for fixed (batch, dims, dtype, k+skip): static input/output buffers + a CUDAGraph."""

    def __init__(self, bank: "RetrievalBank", img: torch.Tensor, txt: Optional[torch.Tensor], kk: int):
        self.img = torch.empty_like(img)
        self.txt = None if txt is None else torch.empty_like(txt)
        self.img.copy_(img)
        if txt is not None:
            self.txt.copy_(txt)
        side = torch.cuda.Stream(device=bank.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up outside capture (allocations, NCCL communicator)
            for _ in range(2):
                bank.search_embeddings(self.img, self.txt, kk=kk)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = bank.search_embeddings(self.img, self.txt, kk=kk)

    def run(self, img: torch.Tensor, txt: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        self.img.copy_(img, non_blocking=True)
        if txt is not None:
            self.txt.copy_(txt, non_blocking=True)
        self.graph.replay()
        return self.out


class RetrievalBank:
    def __init__(self, clip_model=None, clip_tokenize=None, tokenizer=None, device=None, normalise: bool = False,
                 process_group=None, shard: bool = True, max_source_length: int = 512, name: str = "VQADataset",
                 cache_root: str = "cache", additional_root: str = os.path.join("synthetic_data", "cache",
                                                                               "ROCOFeatureDataset"),
                 memoise: bool = True, use_cuda_graph: bool = False, exchange: str = "nccl",
                 fuse_query_cast: bool = True):
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
        self.use_cuda_graph = bool(use_cuda_graph)
        self.fuse_query_cast = bool(fuse_query_cast)
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' (all-gather + merge) or 'p2p' (peer-memory push + flag wait)")
        self.exchange_mode = exchange
        self._p2p: Optional[P2PExchange] = None
        self._process_group = process_group if shard else None
        self._graphs: Dict[tuple, "_SearchGraph"] = {}
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
        self._workspace: Optional[torch.Tensor] = None
        self._memo = None
        K.handle(self.device.index)   # fail loudly now if the library / device is unusable

    # ------------------------------------------------------------------------------------------ bank build
    def create_retrieval_dataset(self, data_loader, prefix, is_training_phase: bool = True, retrieval_k: int = 15,
                                 use_additional_data: bool = False) -> None:
        """Same contract as VQAFeatureDataset.py:118-185 (``prefix`` is accepted and unused, as in the reference)."""
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
            if self.exchange.rank == 0:   # reference-format cache (fp32 [N, 1024] + two pickles), :163-167
                torch.save(torch.cat(host_rows, 0) if host_rows else torch.zeros(0, 0), embedding_path)
                with open(answer_path, "wb") as f:
                    pickle.dump(answers, f)
                with open(question_info_path, "wb") as f:
                    pickle.dump(info, f)

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
        self._graphs = {}

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
        if r == 0:
            self.answer_id.cpu().numpy().tofile(os.path.join(directory, "answer_id.i32"))
            with open(os.path.join(directory, "meta.pkl"), "wb") as f:
                pickle.dump({"answer_strings": self.answer_strings, "info": self.retrieval_question_info}, f)
            manifest = {"format": self.SHARD_FORMAT, "key": key, "n_total": self.n_total, "dim": self.dim,
                        "normalise": self.normalise, "world_size": w,
                        "ranges": [list(shard_bounds(self.n_total, i, w)) for i in range(w)]}
            with open(os.path.join(directory, "manifest.json"), "w") as f:
                json.dump(manifest, f)

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
        self._tables, self._memo, self._graphs = None, None, {}
        return True

    # ------------------------------------------------------------------------------------------ query path
    def _encode(self, batch) -> Tuple[torch.Tensor, torch.Tensor]:
        """The two CLIP calls of VQAFeatureDataset.py:146-147 / :189-190 (stock PyTorch, out of scope)."""
        img = self.clip_model.encode_image(batch["image"].to(self.device, non_blocking=True))
        tokens = self.clip_tokenize(batch["question"]) if self.clip_tokenize is not None else batch["question"]
        if isinstance(tokens, torch.Tensor):
            tokens = tokens.to(self.device, non_blocking=True)
        txt = self.clip_model.encode_text(tokens)
        return img.detach().contiguous(), (None if txt is None else txt.detach().contiguous())

    def search_embeddings(self, image_half: torch.Tensor, text_half: Optional[torch.Tensor] = None, kk: Optional[int] = None
                          ) -> Dict[str, torch.Tensor]:
        """Query halves (device tensors) -> global top-(k+skip): ``score`` fp32 / ``idx`` int32 ``[B, kk]`` and
        ``q_bias`` fp32 ``[B]`` (= -0.5*|q|^2).  Kernel 1 (queries; fused into kernel 2 when D <= 512) -> kernel 2 (+4) -> [all-gather -> kernel 4]."""
        if kk is None:
            kk = self.retrieval_k + (1 if self.is_training_phase else 0)
        b = image_half.shape[0]
        n_local = self.retrieval_embeddings.shape[0]
        fused = self.fuse_query_cast and n_local > 0 and K.search_fused_supported(self.dim, self.device.index)
        if n_local > 0:
            need = K.search_workspace_bytes(b, n_local, self.dim, kk, self.device.index)
            if self._workspace is None or self._workspace.numel() < need:
                self._workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=self.device)
        if fused:
            # D <= 512: concat + (normalise) + bf16 cast happen inside the scan kernel's q-tile load (N3)
            keys, score, idx, qbias = K.search_topk_fused(image_half, text_half, self.retrieval_embeddings, self.bias, kk,
                                                          normalise=self.normalise, idx_base=self.row_begin,
                                                          workspace=self._workspace)
        else:
            q, qbias = K.bank_build(image_half, text_half, normalise=self.normalise)
            if n_local > 0:
                keys, score, idx = K.search_topk(q, self.retrieval_embeddings, self.bias, kk, idx_base=self.row_begin,
                                                 workspace=self._workspace)
            else:
                keys = torch.zeros((b, kk), dtype=torch.int64, device=self.device)
                score = torch.full((b, kk), float("-inf"), device=self.device)
                idx = torch.full((b, kk), -1, dtype=torch.int32, device=self.device)
        if self.exchange.world_size > 1:
            if self.exchange_mode == "p2p":
                if self._p2p is None or self._p2p.cap < b * kk:
                    self._p2p = P2PExchange(self.device, max(b * kk, 4096), self._process_group)
                keys, score, idx = self._p2p.exchange(keys)
            else:
                keys, score, idx = K.merge_topk(self.exchange.gather(keys))
        return {"keys": keys, "score": score, "idx": idx, "q_bias": qbias}      # |q|^2 = -2 * q_bias

    def _graphed_search(self, img: torch.Tensor, txt: Optional[torch.Tensor], kk: int) -> Dict[str, torch.Tensor]:
        """The search chain (kernel 1 -> 2 -> 4 -> [all-gather -> 4]) captured once per shape in a CUDA graph and
        replayed: at small shards the chain is launch-latency-bound.  Every rank of a sharded job must
        take this path for the same shapes (the NCCL all-gather is part of the graph)."""
        key = (tuple(img.shape), img.dtype, None if txt is None else tuple(txt.shape), kk)
        g = self._graphs.get(key)
        if g is None:
            g = _SearchGraph(self, img, txt, kk)
            self._graphs[key] = g
        res = g.run(img, txt)
        if self.memoise:       # the graph's outputs are static buffers; a memoised result must survive the next call
            res = {k_: v.clone() for k_, v in res.items()}
        return res

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

    def _retrieve(self, batch) -> dict:
        """One search per batch object; the reference re-embeds and re-searches the same batch up to 5 times
        (main.py:178-179, 263-270) — results are memoised on the identity of ``batch["image"]``."""
        img_t = batch["image"]
        if self.memoise and self._memo is not None:
            ref, questions, out = self._memo
            if ref() is img_t and questions == list(batch["question"]):
                return out
        skip = 1 if self.is_training_phase else 0
        k = self.retrieval_k
        with torch.no_grad():
            img, txt = self._encode(batch)
        res = self._graphed_search(img, txt, k + skip) if self.use_cuda_graph else \
            self.search_embeddings(img, txt, kk=k + skip)
        out = {"skip": skip, "k": k, "device": res}       # everything stays on the device until somebody asks
        if self.memoise:
            try:
                self._memo = (weakref.ref(img_t), list(batch["question"]), out)
            except TypeError:
                self._memo = None
        return out

    @staticmethod
    def _host(r: dict) -> dict:
        """Single D2H of the search result (replaces the reference's B*k ``tensor.__index__`` syncs at :199)."""
        if "idx" not in r:
            res = r["device"]
            r["idx"] = res["idx"].cpu().numpy()
            r["score"] = res["score"].cpu().numpy()
            r["q_sqnorm"] = res["q_bias"].cpu().numpy() * -2.0
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
        vote = self._vote(r)
        maj, bkt = vote["majority_answer"], vote["bucket"]
        return [prompt_string(int(bkt[i]), self.answer_strings[int(maj[i])], use_quantifier) for i in range(len(maj))]

    def _vote(self, r: dict) -> dict:
        if "vote" not in r:
            res = r["device"]
            dummy = torch.zeros(16, dtype=torch.int32, device=self.device)
            b = res["idx"].shape[0]
            out = K.prompt_gather(res["idx"], r["skip"], self.answer_id, self._lut(r["k"]), dummy,
                                  torch.zeros(b + 1, dtype=torch.int32, device=self.device), dummy,
                                  torch.zeros(9 + len(self.answer_strings), dtype=torch.int32, device=self.device),
                                  True, 0, 1, 2, 1)
            r["vote"] = {k_: out[k_].cpu().numpy() for k_ in ("majority_answer", "majority_count", "bucket", "answer_ids")}
        return r["vote"]

    def retrieve_prompt_ids(self, batch, use_quantifier: bool = True, pad_to: str = "longest"):
        """Additive fast path for ``prepare_input`` (architectures/T5VisionModel.py:143-167): returns the
        ``input_ids`` / ``attention_mask`` the reference's tokenizer call would produce for
        ``task_prefix + question + retrieved_info`` — assembled on the device by kernel 3, no strings involved."""
        r = self._retrieve(batch)
        tables = self._prompt_tables()
        pre_ids, pre_off, longest = tables.prefixes(batch["task"], batch["question"], use_quantifier)
        stride = min(self.max_source_length, longest + tables.tail_bound(use_quantifier))
        out = K.prompt_gather(r["device"]["idx"], r["skip"], self.answer_id, self._lut(r["k"]), pre_ids, pre_off,
                              tables.seg_ids, tables.seg_off, use_quantifier, tables.pad_id, tables.eos_id,
                              self.max_source_length, stride)
        if pad_to == "longest":        # exact shape parity with padding="longest" costs one tiny D2H sync
            longest_out = int(out["length"].max().item())
            return out["input_ids"][:, :longest_out], out["attention_mask"][:, :longest_out]
        return out["input_ids"], out["attention_mask"]
