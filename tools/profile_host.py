"""cProfile of the host side of an end-to-end retrieval step (RetrievalBank.retrieve_prompt_ids_host) on a SMALL bank, so
that the GPU part is short and the Python / ctypes / tokenisation cost per step stands out.  Run on a GPU box:

    python tools/profile_host.py [steps]        -> gpurun_out/host_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from multimodalpromptretrieval_b200 import synthetic as S
from multimodalpromptretrieval_b200.bank import RetrievalBank

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
tok = S.load_tokenizer(os.path.join(ROOT, "tests", "golden", "spm"))
bank = RetrievalBank(tokenizer=tok, device=dev, memoise=False, precomputed_features=True)
n, d, b, k = 200_000, 512, 128, 5
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.randn(n, d, device=dev, generator=g) * 0.44
ans = (np.arange(n) % 24).astype(np.int32)
bank.install_bank([(rows, None)], None, None, is_training_phase=False, retrieval_k=k, answer_ids=ans,
                  answer_strings=S.ROCO_ANSWERS[:24])
q = (torch.randn(b, d) * 0.44).pin_memory()
tasks = [S.TASKS[i % len(S.TASKS)] for i in range(b)]
pool = [{"image": q, "question": [f"{x} #{s}-{i}" for i, x in enumerate(S.make_questions(b, 100 + s))], "task": tasks}
        for s in range(steps + 16)]
for i in range(8):
    bank.retrieve_prompt_ids_host(pool[i])


EPOCH = 32          # the question set of an "epoch": these batches repeat, as a training set does


def run(mode, lo, hi):
    t0 = time.perf_counter()
    if mode == "epoch-two-deep":
        bank.prefetch(pool[lo % EPOCH], True)
        cur = bank.submit_prompt_ids_host(pool[lo % EPOCH])
        bank.prefetch(pool[(lo + 1) % EPOCH], True)
        for i in range(lo, hi):
            bank.prefetch(pool[(i + 2) % EPOCH], True)
            nxt = bank.submit_prompt_ids_host(pool[(i + 1) % EPOCH])
            cur.result()
            cur = nxt
        cur.result()
        return (time.perf_counter() - t0) / (hi - lo) * 1e6
    for i in range(lo, hi):
        if mode == "pipelined":
            bank.prefetch(pool[i + 1], True)
        if mode == "epoch-blocking":
            bank.prefetch(pool[(i + 1) % EPOCH], True)
            bank.retrieve_prompt_ids_host(pool[i % EPOCH])
        else:
            bank.retrieve_prompt_ids_host(pool[i] if mode != "repeated" else pool[i % 4])
    return (time.perf_counter() - t0) / (hi - lo) * 1e6


out = io.StringIO()
for mode in ("sequential", "pipelined", "repeated", "epoch-blocking", "epoch-two-deep"):
    if mode == "pipelined":
        bank.prefetch(pool[8], True)
    if mode == "epoch-blocking":
        for i in range(EPOCH):
            bank.retrieve_prompt_ids_host(pool[i])          # the first epoch tokenises everything once
        bank.prefetch(pool[8 % EPOCH], True)
    us = run(mode, 8, 8 + steps // 2)
    out.write(f"{mode}: {us:.1f} us per end-to-end step (bank {n} x {d}: GPU part ~40 us)\n")
pr = cProfile.Profile()
pr.enable()
if os.environ.get("PROFILE_MODE", "epoch-two-deep") == "epoch-blocking":
    bank.prefetch(pool[(8 + steps // 2) % EPOCH], True)
    run("epoch-blocking", 8 + steps // 2, 8 + steps)
else:
    run("epoch-two-deep", 8 + steps // 2, 8 + steps)
pr.disable()
st = pstats.Stats(pr, stream=out).sort_stats("cumulative")
st.print_stats(45)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "host_profile.txt"), "w").write(out.getvalue())
print(out.getvalue()[:14000])
