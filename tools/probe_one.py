"""One scan shape, a few launches (profiler target):  python tools/probe_one.py B N D KK"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalpromptretrieval_b200 import kernels as K

b, n, d, kk = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = (torch.randn(b, d, device=dev) * 0.3).to(torch.bfloat16)
bank = (torch.randn(n, d, device=dev) * 0.3).to(torch.bfloat16)
bias = -0.5 * (bank.float() ** 2).sum(1)
ws = K.new_workspace(K.search_workspace_bytes(b, n, d, kk), dev)
ok, os_, oi = K.search_topk(q, bank, bias, kk, workspace=ws)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    K.search_topk(q, bank, bias, kk, workspace=ws, out_keys=ok, out_score=os_, out_idx=oi)
e1.record()
torch.cuda.synchronize()
print(f"b={b} n={n} d={d} kk={kk}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us/step, counters={K.debug_counters()}")
if os.environ.get("MPR_DEBUG_COUNTERS"):
    import numpy as np
    ring = K.debug_launch_ring()
    if len(ring) >= 3:
        dur = [(b_ - a_) / 1e3 for a_, b_ in ring]
        gap = [(ring[i + 1][0] - ring[i][1]) / 1e3 for i in range(len(ring) - 1)]
        print("  launch durations (first instruction -> last instruction), us:", " ".join(f"{x:.1f}" for x in dur[-6:]))
        print("  gaps between consecutive launches, us:", " ".join(f"{x:.1f}" for x in gap[-5:]))
    pl = K.search_plan(b, n, d, kk)
    tl = K.debug_timeline(pl["n_ctas"]) / 1e3          # us
    names = ["entry", "q_ready", "producer_done", "g0_loop_end", "g0_written", "g1_loop_end", "g1_written", "cta_done",
             "past_grid_barrier", "tail_done", "tile1_data", "g0_tile1", "g1_tile1", "g0_tile4", "g1_tile4", "g0_tile0_ready", "t0_bound_ready", "t0_chunk0", "t0_chunk1", "t0_chunk3", "t0_released", "t1_bias_staged", "g0_final_flush_done", "t0_published"]
    for k, nm in enumerate(names):
        col = tl[:, k]
        col = col[col >= 0]
        if col.size:
            print(f"  {nm:18s} min {col.min():8.1f}  p50 {np.median(col):8.1f}  p90 {np.percentile(col, 90):8.1f}  max {col.max():8.1f} us")
