"""Host side of the prompt-token gather (kernel 3): pre-tokenised segment tables, the quantifier bucket LUT and the
per-query prefix tokens.

The reference builds one string per query and runs the T5 tokenizer on it
(/root/reference/architectures/T5VisionModel.py:153-167):

    "Answer the {task} question: " + question + "I believe the answer is {bucket} {answer}"      (quantifier on)
    "Answer the {task} question: " + question + "The most frequent answer is {answer}"            (quantifier off)

There is no space between the question and the retrieved sentence, so the question's last word fuses with
"I" / "The" (e.g. ``lung?I``).  Sentencepiece never forms a piece across a whitespace boundary, so the token sequence
of the whole sentence is the concatenation of

    tokens(prefix + "I" | "The")  ‖  tokens("believe the answer is" | "most frequent answer is")
                                  ‖  tokens(bucket)  ‖  tokens(answer)  ‖  </s>

Only the first term depends on the query text; it is tokenised on the host (it does not depend on retrieval, so it
can overlap the bank scan).  Everything else is gathered on the device from tables built once per bank.
"""
from __future__ import annotations

import itertools
import os
import threading
from typing import List, Sequence, Tuple

import numpy as np
import torch

BUCKETS = ["very unlikely", "unlikely", "maybe", "likely", "very likely", "certainly"]   # VQAFeatureDataset.py:188
HEAD_QUANT, TAIL_QUANT = "I", "believe the answer is"                                    # VQAFeatureDataset.py:228
HEAD_PLAIN, TAIL_PLAIN = "The", "most frequent answer is"                                # VQAFeatureDataset.py:230
SEG_QUANT, SEG_PLAIN, SEG_BUCKET0, SEG_ANSWER0 = 0, 1, 2, 8                              # csrc/prompt_gather.cuh


def bucket_lut(k: int) -> np.ndarray:
    """lut[n_votes*(k+1) + max_count] = int(max_count / n_votes * 5), evaluated in Python float64 exactly as
    VQAFeatureDataset.py:223-226 does (e.g. 3/5*5 = 3.0000000000000004 -> 3)."""
    lut = np.zeros((k + 1, k + 1), dtype=np.uint8)
    for n in range(1, k + 1):
        for m in range(1, n + 1):
            certainty = m / n
            lut[n, m] = int(certainty * (len(BUCKETS) - 1))
    return lut.reshape(-1)


def prompt_string(bucket: int, answer: str, use_quantifier: bool) -> str:
    """The sentence retrieve_closest_qa_pairs returns (VQAFeatureDataset.py:228,230)."""
    if use_quantifier:
        return f"{HEAD_QUANT} {TAIL_QUANT} {BUCKETS[bucket]} {answer}"
    return f"{HEAD_PLAIN} {TAIL_PLAIN} {answer}"


def segment_strings(answer_strings: Sequence[str]) -> List[str]:
    """Segment table order expected by kernel 3: two constant tails, six buckets, then the answers."""
    return [TAIL_QUANT, TAIL_PLAIN] + BUCKETS + list(answer_strings)


def prefix_texts(tasks: Sequence[str], questions: Sequence[str], use_quantifier: bool) -> List[str]:
    """Per-query text tokenised on the host: task prefix + question + the fused first word of the retrieved
    sentence (T5VisionModel.py:153,158)."""
    head = HEAD_QUANT if use_quantifier else HEAD_PLAIN
    return [f"Answer the {t} question: " + q + head for t, q in zip(tasks, questions)]


def _csr(rows: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(rows) + 1, dtype=np.int32)
    np.cumsum(list(map(len, rows)), out=off[1:])
    ids = np.array(list(itertools.chain.from_iterable(rows)), dtype=np.int32)
    return ids, off


def _fast_encoder(tokenizer):
    """Batch encoder ``list[str] -> list[list[int]]`` without special tokens.  HF's ``__call__`` costs ~70 us per
    string; when the tokenizer is backed by a sentencepiece file the processor is used directly (C++ batch call), but
    only after it reproduced the HF tokenizer's output on a probe set — otherwise the HF call is kept."""
    def hf(texts):
        return tokenizer(list(texts), add_special_tokens=False)["input_ids"]

    vocab_file = getattr(tokenizer, "vocab_file", None)
    if not vocab_file or not os.path.exists(vocab_file):
        return hf, None
    try:
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor(model_file=vocab_file)
        probe = ["Answer the Modality question:", "what modality is used to take this image?I", "lung?The", "",
                 "very likely", "x-ray, mri", "  two  spaces ", "Mixed Case: 12 3?", TAIL_QUANT, TAIL_PLAIN]
        if sp.encode(probe) == hf(probe):
            return (lambda texts: sp.encode(list(texts))), sp
    except Exception:
        pass
    return hf, None


class PromptTables:
    """Device-resident CSR of the constant segments, the six buckets and every distinct answer of the bank, plus the
    host-side tokenisation of the per-query prefixes."""

    def __init__(self, tokenizer, answer_strings: Sequence[str], device: torch.device):
        self.tokenizer = tokenizer
        self.pad_id = int(tokenizer.pad_token_id)
        self.eos_id = int(tokenizer.eos_token_id)
        self.encode, self._sp = _fast_encoder(tokenizer)
        segs = segment_strings(answer_strings)
        enc = tokenizer(segs, add_special_tokens=False)["input_ids"] if segs else []
        ids, off = _csr(enc)
        self.seg_lens = np.diff(off)
        self.max_answer_len = int(self.seg_lens[SEG_ANSWER0:].max()) if len(answer_strings) else 0
        self.seg_ids = torch.from_numpy(ids if ids.size else np.zeros(1, np.int32)).to(device)
        self.seg_off = torch.from_numpy(off).to(device)
        self.device = device
        self._task_head = {}            # task -> tokens("Answer the {task} question:")
        self._word_cache = {}           # space-separated ASCII chunk -> its tokens (see encode_by_words)
        self._starts_word = None
        self._stage = None              # pinned host staging + device buffers for the prefix CSR (grown on demand)
        self._stage_done = None
        self._native_cache = None       # csrc/token_cache.cpp: the per-batch assembly outside the interpreter
        self._task_index = {}           # task -> row of the head table
        self._head_table = None         # (ids, off) CSR of the task heads
        self._native_lock = threading.Lock()
        self.use_native_cache = True

    def tail_bound(self, use_quantifier: bool) -> int:
        """Upper bound on tokens appended after the prefix (const + bucket + answer + </s>)."""
        if use_quantifier:
            return int(self.seg_lens[SEG_QUANT]) + int(self.seg_lens[SEG_BUCKET0:SEG_ANSWER0].max()) + \
                self.max_answer_len + 1
        return int(self.seg_lens[SEG_PLAIN]) + self.max_answer_len + 1

    def _encode_chunks(self, words: List[str]) -> None:
        """Fills the chunk cache.  Printable chunks are tokenised in ONE sentencepiece call on their space-joined string
        (the per-string overhead of the tokenizer, ~6 us, dominates for short chunks) and split back at the pieces
        that start with the word-boundary marker; anything unusual is tokenised chunk by chunk."""
        cache = self._word_cache
        simple = [w for w in words if w.isprintable()] if self._sp is not None else []
        if simple:
            if self._starts_word is None:      # piece id -> "begins with the word-boundary marker" (built once)
                self._starts_word = np.array([self._sp.id_to_piece(i)[:1] == "\u2581"
                                              for i in range(self._sp.get_piece_size())], dtype=bool)
            ids = np.asarray(self._sp.encode(" ".join(simple)), dtype=np.int64)
            cuts = np.flatnonzero(self._starts_word[ids]) if ids.size else np.zeros(0, dtype=np.int64)
            if cuts.size == len(simple) and cuts[0] == 0:
                bounds = cuts.tolist() + [int(ids.size)]
                flat = ids.tolist()
                for n, w in enumerate(simple):
                    cache[w] = tuple(flat[bounds[n]:bounds[n + 1]])
            else:
                simple = []          # never seen; be safe and fall through to the chunk-by-chunk path
        done = set(simple)
        rest = [w for w in words if w not in done]
        if rest:
            for w, g in zip(rest, self.encode(rest)):
                cache[w] = tuple(g)

    def encode_by_words(self, texts: Sequence[str]) -> List[List[int]]:
        """Tokenises ``texts`` chunk by chunk with a cache.  Sentencepiece never forms a piece across a space, so the
        tokens of a sentence are the concatenation of the tokens of its space-separated chunks (the same property
        kernel 3 relies on); question words recur across batches and epochs, so after warm-up only unseen chunks reach
        the tokenizer (~1.4 ms -> ~0.3 ms per 128 questions).  Restricted to ASCII text, where the tokenizer's NFKC
        normalisation cannot move anything across a space; everything else takes the direct path."""
        cache = self._word_cache
        if len(cache) > 2_000_000:
            cache.clear()
        chunks = [t.split(" ") if t.isascii() else None for t in texts]
        missing = list({w for ws in chunks if ws is not None for w in ws if w and w not in cache})
        if missing:
            self._encode_chunks(missing)
        direct = [i for i, ws in enumerate(chunks) if ws is None]
        direct_ids = dict(zip(direct, self.encode([texts[i] for i in direct]))) if direct else {}
        out: List[List[int]] = []
        for i, ws in enumerate(chunks):
            if ws is None:
                out.append(list(direct_ids[i]))
            else:
                row: List[int] = []
                for w in ws:
                    if w:
                        row.extend(cache[w])
                out.append(row)
        return out

    def _task_heads(self, tasks: Sequence[str]) -> None:
        missing = [t for t in set(tasks) if t not in self._task_head]
        if missing:
            for t, ids in zip(missing, self.encode([f"Answer the {t} question:" for t in missing])):
                self._task_head[t] = list(ids)
            # head table of the native assembler: CSR over the distinct tasks seen so far
            self._task_index = {t: i for i, t in enumerate(self._task_head)}
            ids, off = _csr(list(self._task_head.values()))
            self._head_table = (np.ascontiguousarray(ids), np.ascontiguousarray(off))

    def _prefix_tokens_native(self, tasks: Sequence[str], questions: Sequence[str], head: str):
        """The assembly of :meth:`prefix_tokens` in csrc/token_cache.cpp (mpr_token_cache_assemble): Python only joins and
        encodes the strings and tokenises chunks the cache has never seen.  Returns None for non-ASCII text (the
        interpreter path handles it)."""
        import ctypes as C
        from . import _native
        lib = _native.load()
        if self._native_cache is None:
            with self._native_lock:
                if self._native_cache is None:
                    h = C.c_void_p()
                    if lib.mpr_token_cache_create(C.byref(h)) != 0:
                        return None
                    self._native_cache = h
        n = len(questions)
        try:
            # text i = question i + head; one space between texts (a space only ever ENDS a chunk, so the separator that
            # trails each text changes nothing)
            blob = ((head + " ").join(questions) + head).encode("ascii")
        except UnicodeEncodeError:
            return None
        text_off = np.empty(n + 1, dtype=np.int32)
        text_off[0] = 0
        np.cumsum(np.fromiter(map(len, questions), dtype=np.int32, count=n) + (len(head) + 1), out=text_off[1:])
        if n:
            text_off[n] -= 1
        if int(text_off[n]) != len(blob):
            return None
        idx = self._task_index
        head_index = np.fromiter((idx[t] for t in tasks), dtype=np.int32, count=n)
        h_ids, h_off = self._head_table
        cap = len(blob) + int(np.diff(h_off).max()) * n + 2 * n + 16      # a chunk never yields more tokens than bytes + 1
        out_ids = np.empty(cap, dtype=np.int32)
        out_off = np.empty(n + 1, dtype=np.int32)
        max_missing = 4 * n + 64
        missing = np.empty((max_missing, 2), dtype=np.int32)
        n_missing, longest = C.c_int32(0), C.c_int32(0)
        for _ in range(8):
            rc = lib.mpr_token_cache_assemble(self._native_cache, n, blob, text_off.ctypes.data, h_ids.ctypes.data,
                                              h_off.ctypes.data, head_index.ctypes.data, out_ids.ctypes.data, cap,
                                              out_off.ctypes.data, missing.ctypes.data, max_missing, C.byref(n_missing),
                                              C.byref(longest))
            if rc != 0:
                return None
            if n_missing.value == 0:
                return out_ids[:int(out_off[n])], out_off
            # tokenise the unseen chunks (one sentencepiece call, through the interpreter-side cache) and register them
            words = list({blob[s0:s0 + ln].decode("ascii") for s0, ln in missing[:min(n_missing.value, max_missing)].tolist()})
            fresh = [w for w in words if w not in self._word_cache]
            if fresh:
                self._encode_chunks(fresh)
            if lib.mpr_token_cache_size(self._native_cache) > 4_000_000:
                lib.mpr_token_cache_clear(self._native_cache)
            w_ids, w_off = _csr([self._word_cache[w] for w in words])
            wb = [w.encode("ascii") for w in words]
            c_off = np.zeros(len(wb) + 1, dtype=np.int32)
            np.cumsum(np.fromiter(map(len, wb), dtype=np.int32, count=len(wb)), out=c_off[1:])
            lib.mpr_token_cache_put(self._native_cache, len(wb), b"".join(wb), c_off.ctypes.data,
                                    np.ascontiguousarray(w_ids).ctypes.data, np.ascontiguousarray(w_off).ctypes.data)
        return None

    def prefix_tokens(self, tasks: Sequence[str], questions: Sequence[str], use_quantifier: bool
                      ) -> Tuple[np.ndarray, np.ndarray]:
        """Host CSR of tokens("Answer the {task} question: " + question + "I"|"The").  The task part ends at a
        whitespace boundary, so it is tokenised once per distinct task and concatenated with tokens(question + head)."""
        head = HEAD_QUANT if use_quantifier else HEAD_PLAIN
        self._task_heads(tasks)
        if self.use_native_cache and self._sp is not None:
            out = self._prefix_tokens_native(tasks, questions, head)
            if out is not None:
                return out
        tails = self.encode_by_words([q + head for q in questions])
        return _csr([self._task_head[t] + list(tail) for t, tail in zip(tasks, tails)])

    def prefixes(self, tasks: Sequence[str], questions: Sequence[str], use_quantifier: bool
                 ) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """:meth:`prefix_tokens` staged through pinned memory onto the device; also returns the longest prefix."""
        ids, off = self.prefix_tokens(tasks, questions, use_quantifier)
        n_ids, n_off = max(int(ids.size), 1), int(off.size)
        longest = int(np.diff(off).max()) if len(tasks) else 0
        if self._stage is None or self._stage[0].numel() < n_ids or self._stage[1].numel() < n_off:
            cap_ids, cap_off = max(2 * n_ids, 4096), max(2 * n_off, 512)
            self._stage = (torch.empty(cap_ids, dtype=torch.int32).pin_memory(),
                           torch.empty(cap_off, dtype=torch.int32).pin_memory(),
                           torch.empty(cap_ids, dtype=torch.int32, device=self.device),
                           torch.empty(cap_off, dtype=torch.int32, device=self.device))
            self._stage_done = None
        if self._stage_done is not None:
            self._stage_done.synchronize()          # the previous H2D out of the pinned buffers has finished
        h_ids, h_off, d_ids, d_off = self._stage
        h_ids[:ids.size].copy_(torch.from_numpy(ids))
        h_off[:n_off].copy_(torch.from_numpy(off))
        d_ids[:n_ids].copy_(h_ids[:n_ids], non_blocking=True)
        d_off[:n_off].copy_(h_off[:n_off], non_blocking=True)
        self._stage_done = torch.cuda.Event()
        self._stage_done.record()
        return d_ids[:n_ids], d_off[:n_off], longest
