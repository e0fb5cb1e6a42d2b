"""First-contact GPU probe: kernels vs plain torch on small shapes, verbose diagnostics (not a test)."""
import sys, os, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalpromptretrieval_b200 import kernels as K

torch.manual_seed(0)
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)


def check_build(n, d0, d1, dtype, normalise):
    a = (torch.randn(n, d0, device=dev) * 0.3).to(dtype)
    b = (torch.randn(n, d1, device=dev) * 0.3).to(dtype) if d1 else None
    out, bias = K.bank_build(a, b, normalise=normalise)
    torch.cuda.synchronize()
    x = torch.cat([a, b], 1).float() if d1 else a.float()
    if normalise:
        x = x / x.norm(dim=1, keepdim=True)
    ref = x.to(torch.bfloat16)
    nbad = (out != ref).sum().item()
    refbias = -0.5 * (out.float() ** 2).sum(1)
    print(f"build n={n} d0={d0} d1={d1} {dtype} norm={normalise}: mismatching elems={nbad}/{out.numel()} "
          f"max|bias err|={(bias - refbias).abs().max().item():.3e}", flush=True)


def check_scores(b, n, d):
    q = (torch.randn(b, d, device=dev) * 0.3).to(torch.bfloat16)
    bank = (torch.randn(n, d, device=dev) * 0.3).to(torch.bfloat16)
    bias = -0.5 * (bank.float() ** 2).sum(1)
    s = K.debug_scores(q, bank, bias)
    torch.cuda.synchronize()
    ref = q.float() @ bank.float().T + bias[None, :]
    err = (s - ref).abs()
    nan = torch.isnan(s).sum().item()
    print(f"scores b={b} n={n} d={d}: max err={err.nan_to_num(1e9).max().item():.3e} nan={nan}", flush=True)
    if err.nan_to_num(1e9).max().item() > 1e-2:
        bad = (err.nan_to_num(1e9) > 1e-2).nonzero()
        print("  first bad (q,row):", bad[:8].tolist(), flush=True)
        print("  got", s[0, :8].tolist(), "\n  ref", ref[0, :8].tolist(), flush=True)
    return err.nan_to_num(1e9).max().item()


def check_topk(b, n, d, kk):
    q = (torch.randn(b, d, device=dev) * 0.3).to(torch.bfloat16)
    bank = (torch.randn(n, d, device=dev) * 0.3).to(torch.bfloat16)
    bias = -0.5 * (bank.float() ** 2).sum(1)
    keys, score, idx = K.search_topk(q, bank, bias, kk)
    torch.cuda.synchronize()
    ref = q.float() @ bank.float().T + bias[None, :]
    rs, ri = torch.topk(ref, min(kk, n), dim=1)
    kq = min(kk, n)
    same = (idx[:, :kq].long() == ri).float().mean().item()
    serr = (score[:, :kq] - rs).abs().max().item()
    print(f"topk b={b} n={n} d={d} kk={kk}: idx match={same:.6f} max score err={serr:.3e}", flush=True)


t0 = time.time()
QUICK = bool(os.environ.get("PROBE_AB"))
if not QUICK:
  check_build(1000, 512, 0, torch.float32, False)
  check_build(1000, 512, 512, torch.float16, False)
  check_build(777, 256, 256, torch.float32, True)
h = K.handle(0)
for (b, n, d) in ([] if QUICK else [(16, 128, 64), (16, 128, 512), (128, 1000, 512), (5, 300, 1024), (16, 3072, 1024), (200, 5000, 512)]):
    e = check_scores(b, n, d)
    code = h.device_error()
    if code:
        print("device error code", code, flush=True)
for (b, n, d, kk) in ([(128, 100000, 512, 32)] if QUICK else [(16, 3072, 1024, 1), (16, 14336, 1024, 2), (128, 100000, 512, 5), (1, 50000, 512, 32),
                      (300, 200000, 512, 16), (64, 1000000, 512, 5)]):
    check_topk(b, n, d, kk)
    code = h.device_error()
    if code:
        print("device error code", code, flush=True)
# quick timing
def timing(b, n, d, kk, iters=20):
    q = (torch.randn(b, d, device=dev) * 0.3).to(torch.bfloat16)
    bank = (torch.randn(n, d, device=dev) * 0.3).to(torch.bfloat16)
    bias = -0.5 * (bank.float() ** 2).sum(1)
    ws = K.new_workspace(K.search_workspace_bytes(b, n, d, kk), dev)
    ok, os_, oi = K.search_topk(q, bank, bias, kk, workspace=ws)
    for _ in range(3):
        K.search_topk(q, bank, bias, kk, workspace=ws, out_keys=ok, out_score=os_, out_idx=oi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.search_topk(q, bank, bias, kk, workspace=ws, out_keys=ok, out_score=os_, out_idx=oi)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    pl = K.search_plan(b, n, d, kk)
    if os.environ.get("MPR_DEBUG_COUNTERS"):
        c = K.debug_counters()
        print("  counters per launch:", {k_: round(v / (iters + 4), 1) for k_, v in c.items()}, flush=True)
    print(f"timing b={b} n={n} d={d} kk={kk}: {ms*1e3:.1f} us/scan  {n*d*2/ms/1e6:.1f} GB/s  "
          f"{2*b*n*d/ms/1e9:.1f} TFLOP/s  {b/ms*1e3:.0f} q/s plan={pl}", flush=True)

import os
CASES = [(16, 1250000, 512, 5), (128, 1250000, 512, 5), (256, 1250000, 512, 5), (1024, 1048576, 512, 5),
         (4096, 1048576, 512, 5), (128, 10000000, 512, 5)]
if os.environ.get("PROBE_AB"):
    CASES = [(128, 1250000, 512, 5), (128, 1250000, 512, 16), (128, 1250000, 512, 32), (16, 1250000, 512, 32),
             (128, 10000000, 512, 5), (16, 1000000, 1024, 16)]
if os.environ.get("PROBE_LARGE_K"):
    CASES = [(128, 1250000, 512, 16), (128, 1250000, 512, 32), (16, 1250000, 512, 16), (16, 1000000, 1024, 16)]
if os.environ.get("PROBE_D1024"):
    CASES = [(64, 1000000, 1024, 5), (128, 1000000, 1024, 5), (128, 1000000, 1024, 16), (256, 1000000, 1024, 5),
             (1024, 1000000, 1024, 5), (256, 1000000, 768, 5)]
if os.environ.get("PROBE_FULL"):
    CASES += [(128, 1250000, 512, 1), (128, 1250000, 512, 16), (128, 1250000, 512, 32), (16, 1250000, 512, 32),
              (16, 1000000, 1024, 5), (64, 1000000, 1024, 5), (256, 1000000, 1024, 5), (1, 1250000, 512, 5)]
for (b, n, d, kk) in CASES:
    timing(b, n, d, kk, iters=10 if b > 256 else 20)
print("probe done in", time.time() - t0, "s")
