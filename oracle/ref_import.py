"""Imports the UNMODIFIED reference (/root/reference) under stubs — TEST INFRASTRUCTURE, build container only.

/root/reference does not exist on the GPU box; nothing that runs there may import this module without first
checking ``reference_available()``.  Used by ``oracle/make_golden.py`` (to produce ``tests/golden/``) and by
``tests/test_oracle_golden.py`` (to pin ``oracle/retrieval_oracle.py`` against the live reference).

Why stubs are needed: ``import clip`` (openai-CLIP, not installed), ``import matplotlib``
(utils.py:1,7), ``from ROCO import ...`` (create_mapping.py:10 — needs dataset/ on sys.path); the constructors need
network and datasets, so objects are made with ``__new__`` and the attributes the hot path reads are set by hand.
The functions that are then called — ``VQADataset.retrieve_closest_qa_pairs``, ``VQADataset.create_retrieval_dataset``
and ``T5VisionModel.prepare_input`` — are the reference's own code, byte for byte.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, List, Sequence

import torch

REFERENCE_ROOT = os.environ.get("MPR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dataset", "VQAFeatureDataset.py"))


class _TextTable:
    """Stand-in for ``clip.tokenize``: maps each question string to its (synthetic) 512-d text embedding."""

    def __init__(self):
        self.table: Dict[str, torch.Tensor] = {}

    def __call__(self, questions: Sequence[str]) -> torch.Tensor:
        return torch.stack([self.table[q] for q in questions], 0)


TEXT_TABLE = _TextTable()


class StubClipModel:
    """The "image" tensor of a batch already carries the image-half embedding; "tokenised text" carries the
    text half (see _TextTable).  encode_* are therefore identities — CLIP itself is out of scope."""

    def encode_image(self, x: torch.Tensor) -> torch.Tensor:
        return x

    def encode_text(self, x: torch.Tensor) -> torch.Tensor:
        return x


def _install_stubs() -> None:
    if "clip" not in sys.modules or not getattr(sys.modules["clip"], "_mpr_stub", False):
        clip = types.ModuleType("clip")
        clip._mpr_stub = True
        clip.tokenize = TEXT_TABLE
        clip.load = lambda *a, **k: (StubClipModel(), None)
        sys.modules["clip"] = clip
    for name in ("matplotlib", "matplotlib.patches", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "PIL" not in sys.modules:
        try:
            import PIL  # noqa: F401
        except Exception:
            pil = types.ModuleType("PIL")
            pil.Image = types.ModuleType("PIL.Image")
            sys.modules["PIL"] = pil
            sys.modules["PIL.Image"] = pil.Image
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "dataset")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_reference():
    """Returns (VQADataset class, T5VisionModel class) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    from dataset.VQAFeatureDataset import VQADataset   # type: ignore
    from architectures.T5VisionModel import T5VisionModel   # type: ignore
    return VQADataset, T5VisionModel


def make_reference_dataset(bank: torch.Tensor, answers: List[str], info: Dict[str, List[str]], k: int,
                           is_training_phase: bool):
    """A reference ``VQADataset`` whose retrieval state is set directly (bypassing __init__, which needs data)."""
    VQADataset, _ = load_reference()
    ds = VQADataset.__new__(VQADataset)
    ds.device = "cpu"
    ds.clip_model = StubClipModel()
    ds.retrieval_embeddings = bank.float()
    ds.retrieval_answers = list(answers)
    ds.retrieval_question_info = {key: list(v) for key, v in info.items()}
    ds.is_training_phase = is_training_phase
    ds.retrieval_k = k
    return ds


def make_batch(q_img: torch.Tensor, q_txt: torch.Tensor, questions: Sequence[str], tasks: Sequence[str]) -> dict:
    """A collated batch in the reference's layout (VQAFeatureDataset.py:288-301); registers the text halves."""
    for q, e in zip(questions, q_txt):
        TEXT_TABLE.table[q] = e
    return {"image": q_img, "question": list(questions), "task": list(tasks)}


class _Visual:
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        return torch.zeros(images.shape[0], 50, 512)


class _VisionModel:
    visual = _Visual()


class _T5Stub:
    def __init__(self, vocab: int):
        g = torch.Generator().manual_seed(0)
        self.shared = torch.nn.Embedding(vocab, 512)
        with torch.no_grad():
            self.shared.weight.copy_(torch.randn(vocab, 512, generator=g))


def reference_prepare_input(tokenizer, retrieval_function, batch: dict, use_quantifier: bool,
                            max_source_length: int = 512):
    """Runs the reference's own ``T5VisionModel.prepare_input`` (architectures/T5VisionModel.py:141-184) on a stub
    ``self``; returns ``encoding.input_ids``, ``encoding.attention_mask``."""
    _, T5VisionModel = load_reference()
    self = types.SimpleNamespace(
        retrieval_function=retrieval_function, use_quantifier=use_quantifier, vision_model=_VisionModel(),
        device="cpu", tokenizer=tokenizer, max_source_length=max_source_length, T5_model=_T5Stub(len(tokenizer)),
        use_image_info=True)
    _, _, encoding = T5VisionModel.prepare_input(self, batch)
    return encoding["input_ids"], encoding["attention_mask"]
