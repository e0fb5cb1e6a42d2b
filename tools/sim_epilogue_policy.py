"""CPU model of the scan epilogue's list maintenance (no GPU needed): how many candidates are admitted and how many
SIMT insert iterations a warp executes, as a function of k+s, the pending-buffer depth and the flush rule.

One warp = 32 independent queries (lanes) seeing the same stream of rows; a lane admits a score when it beats its
(possibly stale) threshold; admitted scores wait in the lane's pending buffer until a flush folds EVERY lane's pending
candidates into its sorted list (the warp then executes max-over-lanes insert iterations).  Scores are i.i.d. here,
which is the worst case for insert counts (real banks are no better ordered).

    python tools/sim_epilogue_policy.py
"""
import numpy as np


def simulate(kk, rows, cap, flush_at, groups=1, lanes=32, seed=0, private=False):
    rng = np.random.default_rng(seed)
    res = dict(admitted=0, flushes=0, insert_iters=0, slow_groups=0, ideal=0)
    for g in range(groups):
        n = rows // groups
        s = rng.standard_normal((lanes, n)).astype(np.float32)
        lists = np.full((lanes, kk), -np.inf, dtype=np.float32)
        thr = np.full(lanes, -np.inf, dtype=np.float32)
        pend = [[] for _ in range(lanes)]
        for c0 in range(0, n, 8):
            blk = s[:, c0:c0 + 8]
            hit = blk > thr[:, None]
            if not hit.any():
                continue
            res["slow_groups"] += 1
            for l in np.nonzero(hit.any(1))[0]:
                pend[l].extend(blk[l, hit[l]].tolist())
            lens = np.array([len(p) for p in pend])
            res["admitted"] += int(hit.sum())
            if private:
                todo = np.nonzero(lens > flush_at)[0]
                if len(todo):
                    res["flushes"] += 1
                    res["insert_iters"] += int(lens[todo].max())
                    for l in todo:
                        merged = np.sort(np.concatenate([lists[l], np.array(pend[l], dtype=np.float32)]))[::-1][:kk]
                        lists[l] = merged
                        thr[l] = merged[-1]
                        pend[l] = []
            elif (lens > flush_at).any():
                res["flushes"] += 1
                res["insert_iters"] += int(lens.max())
                for l in range(lanes):
                    if pend[l]:
                        merged = np.sort(np.concatenate([lists[l], np.array(pend[l], dtype=np.float32)]))[::-1][:kk]
                        lists[l] = merged
                        thr[l] = merged[-1]
                        pend[l] = []
        res["ideal"] += int(lanes * kk * (1 + np.log(n / kk)))
    return res


if __name__ == "__main__":
    rows = 8445          # rows per CTA for a 1.25 M-row shard over 148 CTAs
    print("kk cap flush_at groups private | admitted (ideal) flushes insert_iters slow_groups | est. cycles (insert_iter = kk*7+20, slow group = 60)")
    for kk in (5, 16, 32):
        for (cap, fa, groups, private) in ((16, 8, 1, False), (16, 8, 2, False), (32, 24, 1, False), (12, 4, 1, False),
                                           (10, 2, 1, False), (16, 8, 1, True), (16, 4, 1, True), (16, 0, 1, True)):
            r = simulate(kk, rows, cap, fa, groups, private=private)
            cyc = r["insert_iters"] * (kk * 7 + 20) + r["slow_groups"] * 60
            print(f"{kk:2d} {cap:3d} {fa:3d} {groups} {int(private)} | {r['admitted']:6d} ({r['ideal']:5d}) {r['flushes']:5d} "
                  f"{r['insert_iters']:6d} {r['slow_groups']:5d} | {cyc / 1e3:7.0f}k")
