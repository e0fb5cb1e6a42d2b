"""Compute-bound regime probe (cfg4: 4096 queries x 1M x 512): a few scans for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalpromptretrieval_b200 import kernels as K
dev = torch.device("cuda:0")
b, n, d, kk = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 1048576, 512, 5
g = torch.Generator(device=dev).manual_seed(1)
q = (torch.randn(b, d, device=dev, generator=g) * 0.44).to(torch.bfloat16)
bank = (torch.randn(n, d, device=dev, generator=g) * 0.44).to(torch.bfloat16)
_, bias = K.bank_build(bank)
ws = K.new_workspace(K.search_workspace_bytes(b, n, d, kk), dev)
for _ in range(3):
    K.search_topk(q, bank, bias, kk, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    K.search_topk(q, bank, bias, kk, workspace=ws)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"b={b}: {ms*1e3:.0f} us  {2*b*n*d/ms/1e9:.0f} TFLOP/s", K.search_plan(b, n, d, kk))
