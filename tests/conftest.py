import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfg1_k1_test", "cfg1_k1_train", "k5_train", "k15_test_d1024", "k5_dups_test"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _bf16_from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


class GoldenCase:
    """One reference-generated fixture (see oracle/make_golden.py)."""

    def __init__(self, name: str):
        with open(os.path.join(GOLDEN, name + ".json")) as f:
            self.j = json.load(f)
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.name = name
        c = self.j["case"]
        self.k, self.training, self.b = c["k"], c["training"], c["b"]
        self.bank_img = _bf16_from_bits(self.z["bank_img"])
        self.bank_txt = _bf16_from_bits(self.z["bank_txt"])
        self.q_img = _bf16_from_bits(self.z["q_img"])
        self.q_txt = _bf16_from_bits(self.z["q_txt"])
        self.answers = self.j["answers"]
        self.info = self.j["info"]
        self.questions = self.j["questions"]
        self.tasks = self.j["tasks"]

    def bank(self) -> torch.Tensor:
        return torch.cat([self.bank_img, self.bank_txt], 1).float()

    def queries(self) -> torch.Tensor:
        return torch.cat([self.q_img, self.q_txt], 1).float()


@pytest.fixture(scope="session")
def golden_cases():
    return {n: GoldenCase(n) for n in GOLDEN_CASES}


@pytest.fixture(scope="session")
def tokenizer():
    from multimodalpromptretrieval_b200 import synthetic as S
    return S.load_tokenizer(os.path.join(GOLDEN, "spm"))
