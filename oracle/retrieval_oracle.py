"""CPU oracle for the MPR_Gen retrieval hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this module; nothing under ``multimodalpromptretrieval_b200/`` does (a test enforces it).

It is a line-by-line restatement of the reference, using the same third-party ops the reference calls
(torch CPU ``cdist``/``argsort``/``sort``; the algorithm itself lives in PyTorch, pinned ``torch==2.0.1`` in
/root/reference/requirements.txt:7, container has 2.11.0):

    retrieve_closest_qa_pairs      /root/reference/dataset/VQAFeatureDataset.py:187-246
    extend_with_additional_data    /root/reference/dataset/VQAFeatureDataset.py:169-181  (with the :181 bug fixed, see below)
    input_sentences / tokenize     /root/reference/architectures/T5VisionModel.py:153-167

Pinning status: the reference ships NO tests, golden vectors or fixtures for this path, so the
oracle is pinned against *outputs of the reference itself run in the build container*: ``oracle/make_golden.py``
imports the unmodified reference functions under stubs (``oracle/ref_import.py``) and writes ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those files, and against the live reference whenever
``/root/reference`` is present.

Documented deviations from the reference as shipped:
  * ties: ``torch.argsort`` is called with ``stable=False`` in the reference, so its order on exactly equal distances
    is implementation-defined.  The oracle uses ``stable=True`` (= lower index first), which is the
    rule the CUDA path implements; parity checks accept any index whose oracle score is within 1e-3 of the oracle's.
  * ``use_additional_retrieval_data`` crashes in the reference (``dict.extend`` at :181); the oracle
    implements the evident intent (per-key extend).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

BUCKETS = ["very unlikely", "unlikely", "maybe", "likely", "very likely", "certainly"]   # VQAFeatureDataset.py:188


# --------------------------------------------------------------------------------------------- distances
def cdist_mm(q: torch.Tensor, bank: torch.Tensor) -> torch.Tensor:
    """Restatement of ``torch.cdist(q, bank)`` (p=2) on its matmul path, which torch takes when N > 25
    (VQAFeatureDataset.py:192): sqrt(clamp_min([-2q, |q|^2, 1] @ [b, 1, |b|^2]^T, 0))."""
    q = q.float()
    bank = bank.float()
    q_norm = q.pow(2).sum(dim=-1, keepdim=True)
    b_norm = bank.pow(2).sum(dim=-1, keepdim=True)
    q_ = torch.cat([q.mul(-2), q_norm, torch.ones_like(q_norm)], dim=-1)
    b_ = torch.cat([bank, torch.ones_like(b_norm), b_norm], dim=-1)
    return q_.matmul(b_.t()).clamp_min_(0).sqrt_()


def distances(q: torch.Tensor, bank: torch.Tensor) -> torch.Tensor:
    """``dist_matrix`` exactly as the reference computes it (VQAFeatureDataset.py:192)."""
    return torch.cdist(q.float(), bank.float())


def scores_f64(q: torch.Tensor, bank: torch.Tensor) -> torch.Tensor:
    """score = <q,b> - 0.5|b|^2 in float64: the quantity the CUDA path ranks by (argmax score == argmin distance);
    d^2 = |q|^2 - 2*score."""
    q = q.double()
    bank = bank.double()
    return q @ bank.t() - 0.5 * bank.pow(2).sum(-1)[None, :]


def topk_indices(dist: torch.Tensor, k: int, is_training_phase: bool) -> torch.Tensor:
    """VQAFeatureDataset.py:194-197 with the build's deterministic tie rule (stable sort = lower index first)."""
    order = torch.argsort(dist, dim=1, stable=True)
    return order[:, 1:1 + k] if is_training_phase else order[:, 0:k]


# --------------------------------------------------------------------------------------------- vote / prompt
def vote(answers_row: Sequence[str]) -> Tuple[str, int, int]:
    """VQAFeatureDataset.py:217-223: counts in rank order; ``max(counts, key=counts.get)`` returns the first key (in
    insertion order = lowest rank of first occurrence) among those with the maximal count."""
    answer_counts: Dict[str, int] = {}
    for answer in answers_row:
        if answer not in answer_counts:
            answer_counts[answer] = 0
        answer_counts[answer] += 1
    pred_answer = max(answer_counts, key=answer_counts.get)
    return pred_answer, max(answer_counts.values()), sum(answer_counts.values())


def bucket_index(max_count: int, n_votes: int) -> int:
    """VQAFeatureDataset.py:223,226: int(certainty * (len(buckets) - 1)) in Python float64."""
    certainty = max_count / n_votes
    return int(certainty * (len(BUCKETS) - 1))


def prompt_sentence(answers_row: Sequence[str], use_quantifier: bool) -> str:
    """VQAFeatureDataset.py:216-230."""
    pred_answer, max_count, n_votes = vote(answers_row)
    prompt = BUCKETS[bucket_index(max_count, n_votes)]
    if use_quantifier:
        return f"I believe the answer is {prompt} {pred_answer}"
    return f"The most frequent answer is {pred_answer}"


def retrieve_closest_qa_pairs(combined: torch.Tensor, retrieval_embeddings: torch.Tensor,
                              retrieval_answers: Sequence[str], retrieval_question_info: Dict[str, Sequence[str]],
                              retrieval_k: int, is_training_phase: bool, return_ans: bool = False,
                              return_info: Optional[Sequence[str]] = None, return_dists: bool = False,
                              use_quantifier: bool = True):
    """VQAFeatureDataset.py:187-246 after the CLIP encode (``combined`` = cat([image, text], 1).float())."""
    dist_matrix = distances(combined, retrieval_embeddings)                                   # :192
    top_idx = topk_indices(dist_matrix, retrieval_k, is_training_phase)                       # :194-197
    answers = [[retrieval_answers[int(x)] for x in top_idx[i, :]] for i in range(len(top_idx))]   # :199
    retrieved_question_info = []
    if return_info:                                                                           # :202-210
        for indices in top_idx:
            info: List[str] = []
            for idx in indices:
                info.extend(retrieval_question_info[entry][int(idx)] for entry in return_info)
            retrieved_question_info.append(info)
    prompts = [prompt_sentence(row, use_quantifier) for row in answers]                       # :215-230
    if return_ans:                                                                            # :238-246
        return answers
    elif return_info:
        return retrieved_question_info
    elif return_dists:
        smallest = torch.sort(dist_matrix, dim=1).values.detach().cpu().numpy()[:, 0:retrieval_k]
        return list(zip(answers, smallest))
    return prompts


def extend_with_additional_data(embeddings: torch.Tensor, answers: List[str], info: Dict[str, List[str]],
                                extra_embeddings: torch.Tensor, extra_answers: Sequence[str],
                                extra_info: Dict[str, Sequence[str]]):
    """VQAFeatureDataset.py:169-181 with line 181 (``dict.extend``) replaced by the intended per-key extend."""
    embeddings = torch.cat((embeddings, extra_embeddings.float()), dim=0)
    answers = list(answers) + list(extra_answers)
    info = {key: list(vals) + list(extra_info.get(key, [])) for key, vals in info.items()}
    return embeddings, answers, info


# --------------------------------------------------------------------------------------------- prompt splice
def input_sentences(tasks: Sequence[str], questions: Sequence[str], retrieved_info: Sequence[str]) -> List[str]:
    """T5VisionModel.py:153,158 — note: NO space between the question and the retrieved sentence."""
    task_prefixes = [f"Answer the {x} question: " for x in tasks]
    return [task_prefixes[i] + questions[i] + retrieved_info[i] for i in range(len(questions))]


def tokenize(tokenizer, sentences: Sequence[str], max_source_length: int = 512):
    """T5VisionModel.py:161-167."""
    enc = tokenizer(list(sentences), padding="longest", max_length=max_source_length, truncation=True,
                    return_tensors="pt")
    return enc["input_ids"], enc["attention_mask"]


# --------------------------------------------------------------------------------------------- parity helpers
def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """The single host-side rounding both arms consume."""
    return x.to(torch.bfloat16).float()


def check_index_parity(idx_gpu: np.ndarray, score_ref: torch.Tensor, idx_ref: np.ndarray, tol: float = 1e-3
                       ) -> Tuple[int, int]:
    """north_star parity rule.  Position-wise: the CUDA index must equal the oracle's, except where the oracle
    scores of the two candidates differ by <= tol.  Returns (exact matches, tolerated mismatches); raises on a
    violation.  ``score_ref`` is the [B, N] float64 score matrix."""
    assert idx_gpu.shape == idx_ref.shape, (idx_gpu.shape, idx_ref.shape)
    exact = int((idx_gpu == idx_ref).sum())
    tolerated = 0
    bad = np.argwhere(idx_gpu != idx_ref)
    for i, j in bad:
        g, r = int(idx_gpu[i, j]), int(idx_ref[i, j])
        if g < 0 or g >= score_ref.shape[1]:
            raise AssertionError(f"query {i} rank {j}: CUDA index {g} out of range")
        diff = abs(float(score_ref[i, g]) - float(score_ref[i, r]))
        if diff > tol:
            raise AssertionError(f"query {i} rank {j}: CUDA row {g} vs oracle row {r}, score gap {diff:.3e} > {tol}")
        tolerated += 1
    return exact, tolerated


def keys_from(score: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Host restatement of the device key encoding (csrc/topk_key.cuh): hi = order-preserving fp32 image,
    lo = ~row; unsigned compare == (score desc, row asc)."""
    bits = score.astype(np.float32).view(np.uint32).astype(np.uint64)
    ordered = np.where(bits & 0x80000000, (~bits) & 0xFFFFFFFF, bits | 0x80000000)
    low = (~idx.astype(np.uint32)).astype(np.uint64) & 0xFFFFFFFF
    return (ordered << np.uint64(32)) | low


def merge_keys(keys: np.ndarray, kk: int) -> np.ndarray:
    """k-way merge restated on the host: keys [n_lists, B, kk] (u64) -> [B, kk] largest keys, descending."""
    n_lists, b, _ = keys.shape
    flat = np.transpose(keys, (1, 0, 2)).reshape(b, -1)
    out = np.sort(flat, axis=1)[:, ::-1][:, :kk]
    return np.ascontiguousarray(out)


def decode_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(hi & 0x80000000, hi & 0x7FFFFFFF, ~hi).astype(np.uint32)
    score = bits.view(np.float32)
    idx = np.where(keys == 0, -1, (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.int64)).astype(np.int32)
    return score, idx


# --------------------------------------------------------------------------------------------- CPU baseline
def reference_ops_topk(combined: torch.Tensor, retrieval_embeddings: torch.Tensor, retrieval_k: int,
                       is_training_phase: bool) -> torch.Tensor:
    """The reference's scoring + selection exactly as shipped — ``torch.cdist`` then the default (unstable)
    ``torch.argsort`` and the slice (VQAFeatureDataset.py:192-197).  This is what ``bench.py`` times as the CPU
    baseline / ``--impl reference`` arm; it is never used as a checker (ties are implementation-defined)."""
    dist_matrix = torch.cdist(combined.float(), retrieval_embeddings)
    if is_training_phase:
        return torch.argsort(dist_matrix, axis=1)[:, 1:1 + retrieval_k]
    return torch.argsort(dist_matrix, axis=1)[:, 0:retrieval_k]
