"""Small-shape pass over every kernel of the library (compute-sanitizer target: memcheck / racecheck / synccheck, one tool
per run).  Exits non-zero if a result is wrong, so a sanitizer-clean run is also a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpromptretrieval_b200 import _native
from multimodalpromptretrieval_b200 import kernels as K
from multimodalpromptretrieval_b200.sharding import P2PExchange

dev = torch.device("cuda:0")
torch.manual_seed(0)


def ref_topk(q, bank, bias, kk):
    s = q.float() @ bank.float().T + bias[None, :]
    return torch.topk(s, min(kk, bank.shape[0]), dim=1)


def check(name, b, n, d, kk, fused=False, dtype=torch.float32):
    src = torch.randn(n, d, device=dev) * 0.3
    bank, bias = K.bank_build(src)                                   # kernel 1
    raw = (torch.randn(b, d, device=dev) * 0.3).to(dtype)
    q, _ = K.bank_build(raw)
    if fused:
        _, score, idx, _ = K.search_topk_fused(raw, None, bank, bias, kk)
    else:
        _, score, idx = K.search_topk(q, bank, bias, kk)
    rs, ri = ref_topk(q, bank, bias, kk)
    kq = min(kk, n)
    ok = torch.equal(idx[:, :kq].long(), ri) or (score[:, :kq] - rs).abs().max().item() < 1e-3
    print(f"{name}: b={b} n={n} d={d} kk={kk} fused={fused} launches={K.last_launch_count()} ok={ok}", flush=True)
    assert ok, name


check("tmem q-tile, register lists", 16, 3000, 512, 5)
check("tmem q-tile, fused cast", 33, 2000, 256, 6, fused=True, dtype=torch.float16)
check("tmem q-tile, shared-memory lists", 20, 2500, 512, 16)
check("smem q-tile (D=1024)", 16, 1500, 1024, 5)
check("smem q-tile, fused cast", 16, 1500, 1024, 2, fused=True)
check("smem q-tile, CTA pairs + stand-alone tail", 256, 3000, 1024, 3)
check("two launches (grid > one wave)", 1100, 40000, 64, 4)
check("k > N", 4, 3, 64, 5)

# peer-memory exchange with the other rank's delivery pre-populated (world 2, this process = rank 0)
n, d, b, kk, cap = 2000, 128, 8, 3, 256
bank = (torch.randn(n, d, device=dev) * 0.3).to(torch.bfloat16)
_, bias = K.bank_build(bank)
q = (torch.randn(b, d, device=dev) * 0.3).to(torch.bfloat16)
full, _, _ = K.search_topk(q, bank, bias, kk)
x = P2PExchange(dev, cap, world_size=2, rank=0)
other, _, _ = K.search_topk(q, bank[n // 2:].contiguous(), bias[n // 2:].contiguous(), kk, idx_base=n // 2)
host = x.buf.cpu().numpy().copy()
words = host[1024:].view(np.uint64).reshape(4, 2, cap, 2)        # [parity][rank][key][half], tagged with the epoch (1)
okeys = other.cpu().numpy().view(np.uint64).reshape(-1)
tag = np.uint64(1) << np.uint64(32)
words[1, 1, :b * kk, 0] = (okeys & np.uint64(0xFFFFFFFF)) | tag
words[1, 1, :b * kk, 1] = (okeys >> np.uint64(32)) | tag
x.buf.copy_(torch.from_numpy(host))
ws = K.new_workspace(K.search_workspace_bytes(b, n // 2, d, kk), dev)
out = torch.empty((b, kk), dtype=torch.int64, device=dev)
a = _native.RetrieveArgs()
half = bank[:n // 2].contiguous()
hb = bias[:n // 2].contiguous()
a.q_bf16, a.b, a.bank, a.bias, a.n_local, a.d, a.kk = q.data_ptr(), b, half.data_ptr(), hb.data_ptr(), n // 2, d, kk
a.out_keys, a.workspace, a.workspace_bytes = out.data_ptr(), ws.data_ptr(), ws.numel()
x.fill_args(a)
K.retrieve(a, dev)
torch.cuda.synchronize()
assert torch.equal(out, full), "exchange"
print("peer-memory exchange (pre-populated peer): ok", flush=True)

# kernel 4, kernel 3, kernel 5
keys = torch.sort(torch.randint(1, 1 << 62, (4, 9, 5), device=dev), dim=2, descending=True).values.contiguous()
mk, _, _ = K.merge_topk(keys)
assert torch.equal(mk, torch.sort(keys.permute(1, 0, 2).reshape(9, -1), dim=1, descending=True).values[:, :5])
idx = torch.randint(0, 50, (9, 5), dtype=torch.int32, device=dev)
z = torch.zeros(64, dtype=torch.int32, device=dev)
from multimodalpromptretrieval_b200 import prompt as P
K.prompt_gather(idx, 0, torch.randint(0, 7, (50,), dtype=torch.int32, device=dev), torch.from_numpy(P.bucket_lut(5)).to(dev),
                z, z[:10], z, z[:16], True, 0, 1, 8, 4)
table = torch.randn(100, 64, device=dev)
ids = torch.randint(0, 100, (3, 7), device=dev)
emb, _ = K.embed_prompt(ids, torch.ones_like(ids), table, torch.randn(3, 5, 64, device=dev))
assert torch.equal(emb[:, 5:], table[ids])
torch.cuda.synchronize()
assert K.handle(0).device_error() == 0
print("sanitize target done", flush=True)
