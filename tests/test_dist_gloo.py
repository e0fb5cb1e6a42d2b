"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous row shards, the candidate all-gather and the
key-ordered merge reproduce the single-shard top-k exactly.  The per-shard scan and the merge are played by the oracle
here (the CUDA kernels are covered on the GPU by tests/test_gpu_parity.py::test_sharded_search_equals_single_shard)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalpromptretrieval_b200 import sharding
from oracle import retrieval_oracle as O


def _local_topk_keys(q, bank, begin, kk):
    s = O.scores_f64(q, bank).float()
    n = bank.shape[0]
    kq = min(kk, n)
    order = torch.argsort(-s, dim=1, stable=True)[:, :kq]
    score = torch.gather(s, 1, order).numpy()
    keys = np.zeros((q.shape[0], kk), dtype=np.uint64)
    keys[:, :kq] = O.keys_from(score, (order.numpy() + begin).astype(np.int32))
    return keys


def _worker(rank, world, port, n, kk, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        bank = torch.randn(n, 64, generator=g).to(torch.bfloat16).float()
        bank[n // 2 + 3] = bank[2]                      # an exact duplicate across the shard boundary
        q = torch.cat([bank[:4] + 0.01, torch.randn(4, 64, generator=g)]).to(torch.bfloat16).float()
        begin, end = sharding.shard_bounds(n, rank, world)
        ex = sharding.CandidateExchange()
        assert (ex.rank, ex.world_size) == (rank, world)
        local = _local_topk_keys(q, bank[begin:end], begin, kk)
        gathered = ex.gather(torch.from_numpy(local.view(np.int64)).contiguous())
        assert tuple(gathered.shape) == (world, q.shape[0], kk)
        merged = O.merge_keys(gathered.numpy().view(np.uint64), kk)
        full = _local_topk_keys(q, bank, 0, kk)
        ret[rank] = bool(np.array_equal(merged, full))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_gather_merge_equals_single_shard():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 301, 6, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
