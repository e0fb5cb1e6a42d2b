import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalpromptretrieval_b200 import kernels as K
DEV = "cuda:0"
n, b = 30000, 64
for d in (512,):
    g = torch.Generator().manual_seed(9 + d)
    src = torch.randn(n, d, generator=g)
    qsrc = (src[:b] + 0.2 * torch.randn(b, d, generator=g)).contiguous()
    for norm in (True, False):
        bank, bias = K.bank_build(src.to(DEV), normalise=norm)
        q, _ = K.bank_build(qsrc.to(DEV), normalise=norm)
        for kk in (5, 7, 8, 9):
            _, s1, i1 = K.search_topk(q, bank, bias, kk)
            _, s2, i2, _ = K.search_topk_fused(qsrc.to(DEV), None, bank, bias, kk, normalise=norm)
            torch.cuda.synchronize()
            print(f"d={d} norm={norm} kk={kk}: plain idx0={i1[0].tolist()} fused idx0={i2[0].tolist()} "
                  f"n_empty plain={(i1 < 0).sum().item()} fused={(i2 < 0).sum().item()} max|ds|={(s1 - s2).abs().nan_to_num(9).max().item():.2e}")
print("device error", K.handle(0).device_error())
