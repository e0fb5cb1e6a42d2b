"""The config surface and the --train/--test flow of main.py drive the CUDA retrieval path end to end (A12, A13):
retrieval / retrieval_dataset (incl. unions) / retrieval_subset / use_additional_retrieval_data / k / quantifier."""
import json

import pytest

from oracle import retrieval_oracle as O

pytestmark = pytest.mark.gpu


def _cfg(tmp_path, **over):
    cfg = {"seed": 88, "max_source_length": 512, "max_target_length": 128, "dataset": "VQA_RAD",
           "transfer_dataset": "VQA_RAD", "retrieval": 1, "retrieval_dataset": "SLAKE+VQA_RAD", "k": 5, "quantifier": 1,
           "use_additional_retrieval_data": 0, "hyperparameters": {"epochs": 1, "learning_rate": 1e-4, "batch_size": 8},
           "synthetic": {"scale": 0.03, "max_steps": 2}, "cache_root": str(tmp_path / "cache"),
           "synthetic_data_root": str(tmp_path / "synthetic_data")}
    cfg.update(over)
    path = tmp_path / "cfg.json"
    path.write_text(json.dumps(cfg))
    return str(path)


def test_test_flow_with_union_bank_and_analysis_calls(tmp_path):
    from multimodalpromptretrieval_b200.main import main
    rep = main(["--test", "--config", _cfg(tmp_path)])
    assert rep["k"] == 5 and rep["use_quantifier"] is True
    assert rep["bank_rows"] == int(14336 * 0.03) + int(3072 * 0.03)          # SLAKE + VQA_RAD union
    last = rep["last_batch"]
    assert all(len(a) == 5 for a in last["answers"]) and len(last["dists"][0]) == 5
    assert last["prompts"] == [O.prompt_sentence(a, True) for a in last["answers"]]
    assert 0.0 <= rep["gt_in_retrieval"] <= 1.0


def test_train_flow_skip_first_additional_data_no_quantifier(tmp_path):
    from multimodalpromptretrieval_b200.main import main
    path = _cfg(tmp_path, use_additional_retrieval_data=1, quantifier=0, k=3, retrieval_subset=0.5,
                retrieval_dataset="VQA_RAD")
    rep = main(["--train", "--test", "--config", path])
    assert rep["use_quantifier"] is False and len(rep["train_losses"]) == 2
    base = 0
    import random
    from multimodalpromptretrieval_b200.main import load_dataset
    ds = load_dataset("VQA_RAD", "train", 0.03, 88)
    base = len(ds.get_stratified_split(split_fraction=0.5))
    assert rep["bank_rows"] == base + max(64, int(1048576 * 0.03))            # subset + appended ROCO bank
    last = rep["last_batch"]
    assert last["prompts"] == [O.prompt_sentence(a, False) for a in last["answers"]]
    assert all(p.startswith("The most frequent answer is ") for p in last["prompts"])
    del random


def test_k_defaults_to_15_when_missing(tmp_path):
    from multimodalpromptretrieval_b200.main import main
    cfg = json.loads(open(_cfg(tmp_path)).read())
    del cfg["k"], cfg["quantifier"], cfg["retrieval_dataset"]
    (tmp_path / "c2.json").write_text(json.dumps(cfg))
    rep = main(["--test", "--config", str(tmp_path / "c2.json")])
    assert rep["k"] == 15 and rep["use_quantifier"] is True                  # main.py:113-116,127-130
    assert all(len(a) == 15 for a in rep["last_batch"]["answers"])


def test_train_flow_with_device_embeddings(tmp_path):
    """N4 in the flow: prompt ids from the retrieval kernel's tail feed kernel 5 (T5 `shared` gather on the device) and T5
    runs on inputs_embeds; the loss equals the ids-based path's loss (same weights, same batch order), and the gradient
    reaches `shared` through the host wrapper's index_add backward."""
    import torch
    from multimodalpromptretrieval_b200.main import main
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()                    # separate cache roots: both runs build the bank (same RNG consumption)
    rep_ids = main(["--train", "--config", _cfg(tmp_path / "a", device_prompt_embeddings=0)])
    rep_emb = main(["--train", "--config", _cfg(tmp_path / "b", device_prompt_embeddings=1)])
    assert len(rep_emb["train_losses"]) == 2
    assert torch.allclose(torch.tensor(rep_ids["train_losses"][:1]), torch.tensor(rep_emb["train_losses"][:1]), rtol=1e-4)
