"""Same-flag driver for the retrieval path: ``python -m multimodalpromptretrieval_b200.main --train|--test --config C``.

Mirrors the wiring of the reference's ``main.py`` for this path only (/root/reference/main.py:25-37 flags,
:98-130 retrieval config → bank → ``retrieval_function``, :262-294 test-loop analysis calls) and its union-bank
loader (/root/reference/utils.py:89-122).  The training / evaluation logic around it is deliberately minimal: CLIP and
T5 are stock PyTorch (``transformers``), random-initialised because no checkpoints or datasets exist offline, and the
datasets are synthetic stand-ins of the named shapes (``synthetic.py``).  What this file demonstrates — and what
``tests/test_gpu_main_flow.py`` checks — is that the config surface (``retrieval``, ``retrieval_dataset``,
``retrieval_subset``, ``use_additional_retrieval_data``, ``k``, ``quantifier``) and the ``--train/--test`` flow drive the
B200 retrieval kernels exactly as the reference drives its torch ops.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import random
import zlib
from typing import Dict, List

import torch
from torch.utils.data import DataLoader, Dataset

from . import synthetic as S
from .bank import RetrievalBank

# rows / images of the reference's datasets (train split); ROCO = synthetic-QA bank
DATASET_SHAPES = {"SLAKE": (14336, 5000), "VQA_RAD": (3072, 315), "ROCO": (1048576, 80000)}


class SyntheticVQADataset(Dataset):
    """Stand-in for VQASLAKEFeatureDataset / VQARADFeatureDataset / ROCOFeatureDataset: same ``entries`` fields and
    ``__getitem__`` keys (/root/reference/dataset/VQAFeatureDataset.py:288-301).  The "image" is the image id; the
    stand-in CLIP turns ids into embeddings, so no pixel data is needed."""

    def __init__(self, name: str, split: str, scale: float, seed: int):
        rows, images = DATASET_SHAPES[name]
        rows = max(64, int(rows * scale * (1.0 if split == "train" else 0.15)))
        images = max(16, int(images * scale))
        self.name, self.split = name, split
        vocab = S.ROCO_ANSWERS if name == "ROCO" else S.answer_vocab(500, 88)
        g = torch.Generator().manual_seed(seed + sum(map(ord, name + split)))
        img = torch.randint(0, images, (rows,), generator=g).tolist()
        w = 1.0 / torch.arange(1, len(vocab) + 1, dtype=torch.float64) ** 1.2
        ans = torch.multinomial(w / w.sum(), rows, replacement=True, generator=g).tolist()
        qs = S.make_questions(rows, seed + len(name))
        self.entries = [{"image_name": f"{name}_{img[i]}", "question_id": f"{name}_{split}_{i}", "question": qs[i],
                         "answer": vocab[ans[i]], "task": S.TASKS[i % len(S.TASKS)],
                         "question_type": "closed" if vocab[ans[i]] in ("yes", "no") else "open"} for i in range(rows)]
        self.dataroot = f"synthetic://{name}"

    def get_stratified_split(self, split_fraction=0.2, seed=88):      # VQAFeatureDataset.py:249-261
        random.seed(seed)
        by_task: Dict[str, List[int]] = {}
        for i, e in enumerate(self.entries):
            by_task.setdefault(e["task"], []).append(i)
        out: List[int] = []
        for idxs in by_task.values():
            out.extend(random.sample(idxs, int(len(idxs) * split_fraction)))
        return out

    def __len__(self):
        return len(self.entries)

    def __getitem__(self, i):
        e = self.entries[i]
        return {"image": torch.tensor(zlib.crc32(e["image_name"].encode()) % (1 << 31), dtype=torch.int64), "question": e["question"],
                "answer": e["answer"], "task": e["task"], "question_id": e["question_id"],
                "question_type": e["question_type"], "path_to_image": e["image_name"]}


def load_dataset(data_name: str, split: str, scale: float, seed: int) -> SyntheticVQADataset:
    """Union-bank loader with the reference's naming (utils.py:89-122): SLAKE | VQA_RAD | ROCO | COMBINED | "A+B+…".
    As in the reference the result keeps the FIRST dataset's identity (hence its cache directory)."""
    if data_name == "COMBINED":
        data_name = "SLAKE+VQA_RAD"
    names = data_name.split("+")
    combined = None
    for n in names:
        if n == "VQA_RAD" and split == "validate":
            split_n = "train"                                           # utils.py:92-93
        elif n == "ROCO" and split != "train":
            split_n = "test"                                            # utils.py:99-102
        else:
            split_n = split
        d = SyntheticVQADataset(n, split_n, scale, seed)
        if combined is None:
            combined = d
        else:
            combined.entries.extend(d.entries)
    return combined


class HashClip(torch.nn.Module):
    """Stand-in for CLIP ViT-B/32 (out of scope; no weights offline): a frozen random projection of hashed image ids /
    question words to 512-d, scaled to ‖x‖≈7 per half like raw CLIP features.  Same call surface as ``clip``'s model."""

    def __init__(self, dim: int = 512, vocab: int = 4096, seed: int = 88):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.vocab = vocab
        self.img_table = torch.nn.Parameter(torch.randn(vocab, dim, generator=g) * 0.22, requires_grad=False)
        self.txt_table = torch.nn.Parameter(torch.randn(vocab, dim, generator=g) * 0.22, requires_grad=False)

    def encode_image(self, image_ids: torch.Tensor) -> torch.Tensor:
        ids = image_ids.long()
        return self.img_table[ids % self.vocab] + 0.5 * self.img_table[(ids // self.vocab) % self.vocab]

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        emb = self.txt_table[tokens] * (tokens > 0).unsqueeze(-1)
        return emb.sum(1) / 1.5

    def tokenize(self, questions) -> torch.Tensor:
        out = torch.zeros(len(questions), 24, dtype=torch.int64)
        for i, q in enumerate(questions):
            for j, w in enumerate(q.split()[:24]):
                out[i, j] = 1 + sum(ord(c) * (k + 7) for k, c in enumerate(w)) % (self.vocab - 1)
        return out


class T5RetrievalModel(torch.nn.Module):
    """The part of T5VisionModel this path touches (architectures/T5VisionModel.py:141-184,196-234): retrieval →
    prompt ids → T5.  Image tokens are omitted (CLIP token features are out of scope)."""

    def __init__(self, device, tokenizer, retrieval_function=None, retrieval_ids_function=None, use_quantifier=True,
                 max_source_length=512, max_target_length=128, device_embeddings=False):
        super().__init__()
        from transformers import T5Config, T5ForConditionalGeneration
        self.device, self.tokenizer = device, tokenizer
        self.retrieval_function, self.retrieval_ids_function = retrieval_function, retrieval_ids_function
        self.use_quantifier = use_quantifier
        self.max_source_length, self.max_target_length = max_source_length, max_target_length
        self.T5_model = T5ForConditionalGeneration(T5Config(vocab_size=len(tokenizer), decoder_start_token_id=0))
        # N4: ids -> T5_model.shared rows on the device (embed.py / kernel 5) instead of handing ids to T5
        self.device_embeddings = bool(device_embeddings) and retrieval_ids_function is not None

    def _t5_inputs(self, ids, mask):
        if self.device_embeddings:
            from .embed import embed_prompt
            emb, emb_mask = embed_prompt(self.T5_model.shared.weight, ids, mask, None)     # T5VisionModel.py:169,178-180
            return {"inputs_embeds": emb, "attention_mask": emb_mask}
        return {"input_ids": ids, "attention_mask": mask}

    def prepare_input(self, batch):
        if self.retrieval_ids_function is not None:          # device fast path: ids straight from kernel 3
            return self.retrieval_ids_function(batch, use_quantifier=self.use_quantifier)
        if self.retrieval_function:                           # reference path: strings + tokenizer (T5VisionModel.py:143-167)
            info = self.retrieval_function(batch) if self.use_quantifier else \
                self.retrieval_function(batch, use_quantifier=False)
        else:
            info = ["" for _ in batch["task"]]
        sents = [f"Answer the {t} question: " + q + r for t, q, r in zip(batch["task"], batch["question"], info)]
        enc = self.tokenizer(sents, padding="longest", max_length=self.max_source_length, truncation=True,
                             return_tensors="pt")
        return enc["input_ids"].to(self.device), enc["attention_mask"].to(self.device)

    def forward(self, batch):
        ids, mask = self.prepare_input(batch)
        tgt = self.tokenizer(list(batch["answer"]), padding="longest", max_length=self.max_target_length,
                             truncation=True, return_tensors="pt")["input_ids"]
        tgt[tgt == self.tokenizer.pad_token_id] = -100
        return self.T5_model(**self._t5_inputs(ids, mask), labels=tgt.to(self.device)).loss

    @torch.no_grad()
    def predict(self, batch):
        ids, mask = self.prepare_input(batch)
        out = self.T5_model.generate(**self._t5_inputs(ids, mask), do_sample=False, max_new_tokens=8)
        return self.tokenizer.batch_decode(out, skip_special_tokens=True)


def build_retrieval(CFG: dict, args, device, clip_model, tokenizer, dataset_train, scale: float):
    """The reference's retrieval wiring, key for key (/root/reference/main.py:98-130)."""
    if not ("retrieval" in CFG and CFG["retrieval"]):
        return None, None, None
    if "retrieval_dataset" in CFG:
        retrieval_dataset = load_dataset(CFG["retrieval_dataset"], "train", scale, CFG["seed"])
    else:
        retrieval_dataset = dataset_train
    if "retrieval_subset" in CFG:
        split = retrieval_dataset.get_stratified_split(split_fraction=CFG["retrieval_subset"])
        retrieval_dataset.entries = [retrieval_dataset.entries[x] for x in split]
    retrieval_loader = DataLoader(retrieval_dataset, CFG["hyperparameters"]["batch_size"], shuffle=True)
    k = CFG["k"] if "k" in CFG else 15
    additional = bool("use_additional_retrieval_data" in CFG and CFG["use_additional_retrieval_data"])
    cache_root = CFG.get("cache_root", "cache")
    additional_root = os.path.join(CFG.get("synthetic_data_root", "synthetic_data"), "cache", "ROCOFeatureDataset")
    bank = RetrievalBank(clip_model=clip_model, clip_tokenize=clip_model.tokenize, tokenizer=tokenizer, device=device,
                         name=f"{retrieval_dataset.name}FeatureDataset", cache_root=cache_root,
                         additional_root=additional_root, max_source_length=CFG.get("max_source_length", 512))
    if additional:
        print(f"Using {k}-nn retrieval from {retrieval_dataset.dataroot} with additional synthetic data ...")
        if not os.path.exists(os.path.join(additional_root, "embedding.pt")):
            write_additional_cache(additional_root, clip_model, device, scale, CFG["seed"])
    else:
        print(f"Using {k}-nn retrieval from {retrieval_dataset.dataroot} with only training data ...")
    bank.create_retrieval_dataset(retrieval_loader, "prefix", is_training_phase=args.train, retrieval_k=k,
                                  use_additional_data=additional)
    return bank, retrieval_loader, k


@torch.no_grad()
def write_additional_cache(root: str, clip_model, device, scale: float, seed: int) -> None:
    """Produces synthetic_data/cache/ROCOFeatureDataset/{embedding.pt,answers.pkl,answer_types.pkl} — the files the
    reference expects (VQAFeatureDataset.py:170-172) but nothing in it writes."""
    ds = load_dataset("ROCO", "train", scale, seed)
    os.makedirs(root, exist_ok=True)
    embs, answers, info = [], [], {"question_type": [], "question_id": [], "question": []}
    for batch in DataLoader(ds, 512):
        e = torch.cat([clip_model.encode_image(batch["image"].to(device)),
                       clip_model.encode_text(clip_model.tokenize(batch["question"]).to(device))], 1)
        embs.append(e.float().cpu())
        answers.extend(batch["answer"])
        for key in info:
            info[key].extend(batch[key])
    torch.save(torch.cat(embs, 0), os.path.join(root, "embedding.pt"))
    pickle.dump(answers, open(os.path.join(root, "answers.pkl"), "wb"))
    pickle.dump(info, open(os.path.join(root, "answer_types.pkl"), "wb"))


def main(argv=None) -> dict:
    parser = argparse.ArgumentParser()                                 # flags of /root/reference/main.py:25-34
    parser.add_argument("--train", action="store_true", help="train a model")
    parser.add_argument("--resume", action="store_true", help="Resume model training")
    parser.add_argument("--test", action="store_true", help="test a model")
    parser.add_argument("--eval", action="store_true", help="evaluate a model")
    parser.add_argument("--config", help="config file name in the config folder")
    parser.add_argument("--gpu_id", default="0", help="ID of GPU")
    parser.add_argument("--model_file", help="optional path to model to save/load")
    parser.add_argument("--qid", help="Question ID to analyze")
    parser.add_argument("--tokenizer_dir", default=os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "tests", "golden", "spm"), help="directory with a T5 spiece.model")
    args = parser.parse_args(argv)
    CFG = json.load(open(args.config))
    random.seed(CFG["seed"])
    torch.manual_seed(CFG["seed"])
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    syn = CFG.get("synthetic", {})
    scale, max_steps = float(syn.get("scale", 0.02)), int(syn.get("max_steps", 3))
    data_name = CFG["dataset"]
    if "transfer_dataset" in CFG and not args.train:                    # main.py:67-69
        data_name = CFG["transfer_dataset"]
    bs = CFG["hyperparameters"]["batch_size"]
    dataset_train = load_dataset(data_name, "train", scale, CFG["seed"])
    dataset_test = load_dataset(data_name, "test", scale, CFG["seed"])
    train_loader = DataLoader(dataset_train, bs, shuffle=True)
    test_loader = DataLoader(dataset_test, bs, shuffle=True)

    tokenizer = S.load_tokenizer(args.tokenizer_dir)
    clip_model = HashClip().to(device)
    bank, retrieval_loader, k = build_retrieval(CFG, args, device, clip_model, tokenizer, dataset_train, scale)
    retrieval_function = bank.retrieve_closest_qa_pairs if bank else None                     # main.py:123
    use_quantifier = not ("quantifier" in CFG and not CFG["quantifier"])                      # main.py:127-130
    model = T5RetrievalModel(device, tokenizer, retrieval_function=retrieval_function,
                             retrieval_ids_function=bank.retrieve_prompt_ids if bank and CFG.get("device_prompt_ids", 1) else None,
                             use_quantifier=use_quantifier, max_source_length=CFG.get("max_source_length", 512),
                             max_target_length=CFG.get("max_target_length", 128),
                             device_embeddings=CFG.get("device_prompt_embeddings", 0)).to(device)
    report = {"k": k, "use_quantifier": use_quantifier, "bank_rows": bank.n_total if bank else 0}

    if args.train:
        opt = torch.optim.AdamW(model.parameters(), lr=CFG["hyperparameters"]["learning_rate"])
        losses = []
        for step, batch in enumerate(train_loader):
            if step >= max_steps:
                break
            loss = model(batch)                       # retrieval call 1 (forward)        main.py:178
            model.predict(batch)                      # retrieval call 2 (memoised)       main.py:179
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        report["train_losses"] = losses
        print(f"train losses: {losses}")

    if args.test:
        model.eval()
        gt_in_retrieval = total = 0
        type_consistency: List[float] = []
        for step, batch in enumerate(test_loader):
            if step >= max_steps:
                break
            predicted = model.predict(batch)
            if bank:                                                                          # main.py:266-270
                ds = bank
                retrieved_answers = ds.retrieve_closest_qa_pairs(batch, return_ans=True)
                retrieved_types = ds.retrieve_closest_qa_pairs(batch, return_info=["question_type"])
                retrieved_qinfo = ds.retrieve_closest_qa_pairs(batch, return_info=["question", "question_id"])
                retrieved_dists = ds.retrieve_closest_qa_pairs(batch, return_dists=True)
                for i in range(len(predicted)):
                    total += 1
                    gt_in_retrieval += int(batch["answer"][i].lower() in retrieved_answers[i])
                    type_consistency.append(sum(1 for x in retrieved_types[i] if x == batch["question_type"][i]) /
                                            len(retrieved_types[i]))
                report["last_batch"] = {"prompts": ds.retrieve_closest_qa_pairs(batch, use_quantifier=use_quantifier),
                                        "answers": retrieved_answers, "qinfo": retrieved_qinfo[0][:2],
                                        "dists": [d.tolist() for _, d in retrieved_dists][:2]}
        if total:
            report["gt_in_retrieval"] = gt_in_retrieval / total
            report["type_consistency"] = sum(type_consistency) / len(type_consistency)
            print(f"Ground truth in retrieved set: {gt_in_retrieval / total:.3f}  "
                  f"question-type consistency: {report['type_consistency']:.3f}")
    return report


if __name__ == "__main__":
    main()
