"""Generates ``tests/golden/`` by running the UNMODIFIED reference (/root/reference) under stubs.

TEST INFRASTRUCTURE — run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

For every case it stores the synthetic inputs (bf16-exact values as uint16 bit patterns, so the CUDA arm and the
oracle consume identical numbers — SURVEY.md H3) and what the reference's own functions returned for them:

    VQADataset.retrieve_closest_qa_pairs   prompts (quantifier on/off), return_ans, return_info x2, return_dists
    T5VisionModel.prepare_input            encoding.input_ids / attention_mask (synthetic sentencepiece T5 vocab)

Cases without exact duplicate rows are tie-free, so the reference's unstable argsort is deterministic on them and
they pin indices bit-exactly; the ``dups`` case keeps exact duplicates and is compared under the 1e-3 rule.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodalpromptretrieval_b200 import synthetic as S   # noqa: E402
from oracle import ref_import as R                          # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SPM_DIR = os.path.join(GOLDEN, "spm")

CASES = [
    # name, n_rows, n_images, d_half, batch, k, training, dup_frac
    dict(name="cfg1_k1_test", n_rows=768, n_images=200, d_half=64, b=16, k=1, training=False, dup_frac=0.0),
    dict(name="cfg1_k1_train", n_rows=768, n_images=200, d_half=64, b=16, k=1, training=True, dup_frac=0.0),
    dict(name="k5_train", n_rows=1024, n_images=300, d_half=64, b=16, k=5, training=True, dup_frac=0.0),
    dict(name="k15_test_d1024", n_rows=384, n_images=100, d_half=512, b=8, k=15, training=False, dup_frac=0.0),
    dict(name="k5_dups_test", n_rows=1024, n_images=300, d_half=64, b=16, k=5, training=False, dup_frac=0.05),
]


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def build_case(c: dict, tokenizer) -> None:
    seed = 88 + sum(ord(ch) for ch in c["name"])
    vocab = S.answer_vocab(60, 88)     # small vocabulary so that k-NN votes actually collide
    bank = S.make_bank(c["n_rows"], c["n_images"], c["d_half"], c["dup_frac"], seed=seed, answers=vocab)
    qs = S.make_queries(bank, c["b"], seed=seed + 1)
    ds = R.make_reference_dataset(bank.combined(), bank.answers, bank.info, c["k"], c["training"])
    batch = R.make_batch(qs.image_half, qs.text_half, qs.questions, qs.tasks)

    out = {"case": c}
    out["prompts_quant"] = ds.retrieve_closest_qa_pairs(batch)
    out["prompts_plain"] = ds.retrieve_closest_qa_pairs(batch, use_quantifier=False)
    out["return_ans"] = ds.retrieve_closest_qa_pairs(batch, return_ans=True)
    out["return_info_type"] = ds.retrieve_closest_qa_pairs(batch, return_info=["question_type"])
    out["return_info_q_id"] = ds.retrieve_closest_qa_pairs(batch, return_info=["question", "question_id"])
    dists = ds.retrieve_closest_qa_pairs(batch, return_dists=True)
    out["return_dists_answers"] = [list(a) for a, _ in dists]
    dist_arr = np.stack([d for _, d in dists]).astype(np.float32)
    ids_q, mask_q = R.reference_prepare_input(tokenizer, ds.retrieve_closest_qa_pairs, batch, use_quantifier=True)
    ids_p, mask_p = R.reference_prepare_input(tokenizer, ds.retrieve_closest_qa_pairs, batch, use_quantifier=False)
    # the reference's own top-k indices (argsort slice), for index-level parity
    dm = torch.cdist(qs.combined().float(), bank.combined().float())
    order = torch.argsort(dm, axis=1)
    top = order[:, 1:1 + c["k"]] if c["training"] else order[:, 0:c["k"]]

    out["answers"] = bank.answers
    out["info"] = bank.info
    out["questions"] = qs.questions
    out["tasks"] = qs.tasks
    with open(os.path.join(GOLDEN, c["name"] + ".json"), "w") as f:
        json.dump(out, f)
    np.savez_compressed(
        os.path.join(GOLDEN, c["name"] + ".npz"),
        bank_img=bf16_bits(bank.image_half), bank_txt=bf16_bits(bank.text_half),
        q_img=bf16_bits(qs.image_half), q_txt=bf16_bits(qs.text_half),
        return_dists=dist_arr, top_idx=top.numpy().astype(np.int32),
        input_ids_quant=ids_q.numpy(), attention_mask_quant=mask_q.numpy(),
        input_ids_plain=ids_p.numpy(), attention_mask_plain=mask_p.numpy())
    print(f"{c['name']}: N={bank.n} D={2 * c['d_half']} B={c['b']} k={c['k']} train={c['training']} "
          f"L={ids_q.shape[1]} e.g. {out['prompts_quant'][0]!r}")


def main() -> None:
    if not R.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(GOLDEN, exist_ok=True)
    if not os.path.exists(os.path.join(SPM_DIR, "spiece.model")):
        S.train_tokenizer(SPM_DIR, vocab_size=1000)
    tokenizer = S.load_tokenizer(SPM_DIR)
    torch.manual_seed(88)
    for c in CASES:
        build_case(c, tokenizer)


if __name__ == "__main__":
    main()
