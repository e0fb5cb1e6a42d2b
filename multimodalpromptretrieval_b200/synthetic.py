"""Synthetic stand-ins for the datasets, CLIP embeddings and T5 vocabulary that are not available offline.

Shapes follow SURVEY.md §8(d): SLAKE/VQA_RAD-like banks (rows = QA pairs, several QAs share an image so their image
halves are identical; a few exact duplicate rows as VQA_RAD produces, /root/reference/dataset/VQA_RAD.py:37-50) and the
ROCO synthetic-QA bank (24 distinct answers, /root/reference/synthetic_data/generate_roco_questions.py).  Embeddings
are scaled so that ‖row‖ ≈ 10 like raw CLIP ViT-B/32 features.  Everything is seeded (default 88 = the reference's
``config/experiment.json:2``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

TASKS = ["Modality", "Plane", "Organ", "Abnormality", "Position", "Color", "Size", "Shape", "Quantity", "KG"]

_ORGANS = ["lung", "liver", "heart", "brain", "kidney", "spleen", "bladder", "stomach", "colon", "pancreas",
           "esophagus", "trachea", "spinal cord", "rectum", "small bowel", "gallbladder", "uterus", "prostate"]
_SIDES = ["left", "right", "upper", "lower", "bilateral", "center", "upper left", "lower right", "top", "bottom"]
_MODAL = ["ct", "mri", "x-ray", "t1", "t2", "ultrasound", "pet", "angiogram"]
_PLANES = ["axial", "coronal", "sagittal", "transverse plane", "coronal plane"]
_DISEASE = ["pneumonia", "cardiomegaly", "nodule", "brain edema", "brain tumor", "atelectasis", "pneumothorax",
            "pulmonary mass", "liver cancer", "effusion", "infiltration", "lung cancer", "fracture", "hemorrhage"]
_MISC = ["yes", "no", "1", "2", "3", "4", "5", "black", "white", "gray", "hyperdense", "hypodense", "circular",
         "oval", "irregular", "almost the same", "much", "not seen", "head", "chest", "abdomen", "neck", "pelvic cavity"]

ROCO_ANSWERS = ["yes", "no", "musculoskeletal", "cardiovascular", "respiratory", "digestive", "nervous", "urinary",
                "lung", "heart", "brain", "liver", "kidney", "ct", "mri", "x-ray", "ultrasound", "pet", "angiography",
                "mammography", "axial", "coronal", "sagittal", "transverse"]

_Q_TEMPLATES = [
    "what modality is used to take this image?", "which part of the body does this image belong to?",
    "does the picture contain {o}?", "where is the {o} in this image?", "is the {o} healthy?",
    "what is the main organ in the image?", "which organ is abnormal, {o} or {o2}?", "what diseases are included in the picture?",
    "is this a {m} scan?", "what color is the {o} in the picture?", "how many {o}s are there in this image?",
    "what is the scanning plane of this image?", "what is the shape of the {o}?", "is there {d} in the {o}?",
    "where is the {d} located?", "which is bigger in this image, {o} or {o2}?", "what is the function of the {o}?",
]


def answer_vocab(n: int = 500, seed: int = 88) -> List[str]:
    """``n`` distinct lower-case answer strings (the reference lower-cases answers, VQAFeatureDataset.py:72)."""
    g = torch.Generator().manual_seed(seed)
    base = list(dict.fromkeys(_MISC + _ORGANS + _MODAL + _PLANES + _DISEASE + _SIDES))
    out = list(base)
    pools = [_SIDES, _ORGANS, _DISEASE]
    while len(out) < n:
        i = torch.randint(0, 1 << 30, (4,), generator=g).tolist()
        a, b, c = _SIDES[i[0] % len(_SIDES)], _ORGANS[i[1] % len(_ORGANS)], _DISEASE[i[2] % len(_DISEASE)]
        cand = [f"{a} {b}", f"{b}, {c}", f"{c} in the {a} {b}", f"{a} {b} and {_ORGANS[i[3] % len(_ORGANS)]}"][i[3] % 4]
        if cand not in out:
            out.append(cand)
    del pools
    return out[:n]


def make_questions(n: int, seed: int = 88) -> List[str]:
    g = torch.Generator().manual_seed(seed + 1)
    r = torch.randint(0, 1 << 30, (n, 5), generator=g).tolist()
    qs = []
    for a, b, c, d, e in r:
        t = _Q_TEMPLATES[a % len(_Q_TEMPLATES)]
        qs.append(t.format(o=_ORGANS[b % len(_ORGANS)], o2=_ORGANS[c % len(_ORGANS)], m=_MODAL[d % len(_MODAL)],
                           d=_DISEASE[e % len(_DISEASE)]))
    return qs


@dataclass
class SyntheticBank:
    image_half: torch.Tensor          # [n, d_half] fp32 — what clip.encode_image would return
    text_half: torch.Tensor           # [n, d_half] fp32 — what clip.encode_text would return
    answers: List[str]
    info: Dict[str, List[str]]        # question_type / question_id / question

    @property
    def n(self) -> int:
        return self.image_half.shape[0]

    def combined(self) -> torch.Tensor:
        return torch.cat([self.image_half, self.text_half], 1)


def make_bank(n_rows: int, n_images: int, d_half: int = 512, dup_frac: float = 0.02, seed: int = 88,
              answers: Optional[List[str]] = None, zipf: float = 1.2, norm: float = 10.0,
              round_bf16: bool = True, id_prefix: str = "q") -> SyntheticBank:
    """Rows = QA pairs over ``n_images`` images; ``dup_frac`` of the rows are exact copies of earlier rows."""
    g = torch.Generator().manual_seed(seed)
    scale = norm / (2 * d_half) ** 0.5
    img_table = torch.randn(n_images, d_half, generator=g) * scale
    img_of_row = torch.randint(0, n_images, (n_rows,), generator=g)
    image_half = img_table[img_of_row]
    text_half = torch.randn(n_rows, d_half, generator=g) * scale
    n_dup = int(n_rows * dup_frac)
    if n_dup > 0 and n_rows > 1:
        dst = torch.randperm(n_rows - 1, generator=g)[:n_dup] + 1
        src = (torch.rand(n_dup, generator=g) * dst.float()).long()     # an earlier row
        image_half[dst] = image_half[src]
        text_half[dst] = text_half[src]
    if round_bf16:
        image_half = image_half.to(torch.bfloat16).float()
        text_half = text_half.to(torch.bfloat16).float()
    vocab = answers if answers is not None else answer_vocab(500, seed)
    w = 1.0 / torch.arange(1, len(vocab) + 1, dtype=torch.float64) ** zipf
    ans_idx = torch.multinomial(w / w.sum(), n_rows, replacement=True, generator=g).tolist()
    row_answers = [vocab[i] for i in ans_idx]
    questions = make_questions(n_rows, seed)
    qtypes = ["closed" if a in ("yes", "no") else "open" for a in row_answers]
    info = {"question_type": qtypes, "question_id": [f"{id_prefix}{i}" for i in range(n_rows)], "question": questions}
    return SyntheticBank(image_half.contiguous(), text_half.contiguous(), row_answers, info)


@dataclass
class SyntheticQueries:
    image_half: torch.Tensor
    text_half: torch.Tensor
    questions: List[str]
    tasks: List[str]

    def combined(self) -> torch.Tensor:
        return torch.cat([self.image_half, self.text_half], 1)


def make_queries(bank: SyntheticBank, b: int, seed: int = 89, noise: float = 0.1, round_bf16: bool = True
                 ) -> SyntheticQueries:
    """Half the batch = a bank row + small noise (near-self match, exercises the training-phase skip-first);
    the rest unrelated."""
    g = torch.Generator().manual_seed(seed)
    d_half = bank.image_half.shape[1]
    scale = bank.image_half.std().item()
    src = torch.randint(0, bank.n, (b,), generator=g)
    qi = bank.image_half[src] + noise * scale * torch.randn(b, d_half, generator=g)
    qt = bank.text_half[src] + noise * scale * torch.randn(b, d_half, generator=g)
    fresh = torch.arange(b) % 2 == 1
    qi[fresh] = torch.randn(int(fresh.sum()), d_half, generator=g) * scale
    qt[fresh] = torch.randn(int(fresh.sum()), d_half, generator=g) * scale
    if round_bf16:
        qi, qt = qi.to(torch.bfloat16).float(), qt.to(torch.bfloat16).float()
    questions = [f"{q} #{i}" for i, q in enumerate(make_questions(b, seed + 7))]   # unique strings
    tasks = [TASKS[int(i) % len(TASKS)] for i in torch.randint(0, 1 << 20, (b,), generator=g).tolist()]
    return SyntheticQueries(qi.contiguous(), qt.contiguous(), questions, tasks)


# ------------------------------------------------------------------------------------------------ tokenizer
def tokenizer_corpus(seed: int = 88) -> List[str]:
    vocab = answer_vocab(500, seed) + ROCO_ANSWERS
    qs = make_questions(4000, seed)
    lines = list(vocab) + qs
    buckets = ["very unlikely", "unlikely", "maybe", "likely", "very likely", "certainly"]
    for i, q in enumerate(qs[:1500]):
        t = TASKS[i % len(TASKS)]
        a = vocab[i % len(vocab)]
        lines.append(f"Answer the {t} question: {q}I believe the answer is {buckets[i % 6]} {a}")
        lines.append(f"Answer the {t} question: {q}The most frequent answer is {a}")
    return lines


def train_tokenizer(out_dir: str, vocab_size: int = 1000, seed: int = 88) -> str:
    """Trains a small sentencepiece unigram model laid out like t5-small's (pad=0, </s>=1, <unk>=2) into
    ``out_dir/spiece.model``.  There is no T5 vocabulary on disk and no network."""
    import sentencepiece as spm
    os.makedirs(out_dir, exist_ok=True)
    corpus = os.path.join(out_dir, "corpus.txt")
    with open(corpus, "w") as f:
        f.write("\n".join(tokenizer_corpus(seed)))
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(out_dir, "spiece"), vocab_size=vocab_size,
                                   model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1,
                                   character_coverage=1.0, hard_vocab_limit=False, minloglevel=2)
    os.remove(corpus)
    if os.path.exists(os.path.join(out_dir, "spiece.vocab")):
        os.remove(os.path.join(out_dir, "spiece.vocab"))
    return os.path.join(out_dir, "spiece.model")


def load_tokenizer(model_dir: str):
    """T5Tokenizer over a local spiece.model, with the reference's extra ``[itk]`` token
    (/root/reference/architectures/T5VisionModel.py:57-58)."""
    from transformers import T5Tokenizer
    tok = T5Tokenizer.from_pretrained(model_dir)
    tok.add_tokens(["[itk]"])
    return tok
