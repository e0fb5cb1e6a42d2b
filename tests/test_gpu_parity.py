"""GPU parity tests: every kernel is driven through the C ABI (libmpr_b200.so) and checked against the CPU oracle
(oracle/retrieval_oracle.py) on identical bf16-rounded inputs, plus the reference-generated golden fixtures.

Parity bar (BASELINE.json north_star): retrieved indices and prompt token ids bit-exact, except where two candidates'
oracle scores differ by <= 1e-3; scores within 1e-3 absolute.
"""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as O
from tests.conftest import GOLDEN_CASES

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def K():
    from multimodalpromptretrieval_b200 import kernels
    kernels.handle(0)
    return kernels


def dev():
    return torch.device("cuda:0")


def clip_like(n, d, seed, norm=10.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, d, generator=g) * (norm / d ** 0.5)).to(torch.bfloat16)


def oracle_topk(q, bank, kk):
    s = O.scores_f64(q.float(), bank.float())
    order = torch.argsort(-s, dim=1, stable=True)[:, :kk]
    return s, order.numpy(), torch.gather(s, 1, order).numpy()


# ------------------------------------------------------------------------------------------------ kernel 1
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("d0,d1", [(512, 512), (512, 0), (64, 0), (256, 64)])
def test_bank_build_cast_and_bias(K, dtype, d0, d1):
    g = torch.Generator().manual_seed(d0 + d1)
    n = 1037
    a = (torch.randn(n, d0, generator=g) * 0.3).to(dtype)
    b = (torch.randn(n, d1, generator=g) * 0.3).to(dtype) if d1 else None
    out, bias = K.bank_build(a.to(dev()), None if b is None else b.to(dev()))
    x = torch.cat([a, b], 1).float() if d1 else a.float()
    ref = x.to(torch.bfloat16)
    assert torch.equal(out.cpu(), ref)                                   # cast is bit-exact (RN-even)
    ref_bias = -0.5 * ref.double().pow(2).sum(1)
    assert (bias.cpu().double() - ref_bias).abs().max().item() < 1e-4


def test_bank_build_normalise(K):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2000, 1024, generator=g) * 0.3
    x[7] = 0
    out, bias = K.bank_build(x.to(dev()), normalise=True)
    ref = (x / x.norm(dim=1, keepdim=True).clamp_min(1e-30)).to(torch.bfloat16)
    out_c = out.cpu()
    assert torch.equal(out_c[7], torch.zeros(1024, dtype=torch.bfloat16))
    diff_ulps = (out_c.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    assert diff_ulps.max().item() <= 1                                   # norm summation order: at most 1 bf16 ulp
    assert (diff_ulps > 0).float().mean().item() < 1e-3
    assert (bias.cpu()[:7] + 0.5).abs().max().item() < 2e-3              # |row| = 1 up to bf16 rounding


def test_bank_build_rejects_bad_arguments(K):
    from multimodalpromptretrieval_b200._native import NativeError
    with pytest.raises(NativeError):
        K.bank_build(torch.zeros(4, 100, device=dev()))                  # D % 64 != 0
    with pytest.raises(NativeError):
        K.bank_build(torch.zeros(4, 4096, device=dev()))                 # D > 2048


# ------------------------------------------------------------------------------------------------ kernel 2: scores
@pytest.mark.parametrize("b,n,d", [(1, 1, 64), (16, 128, 64), (16, 3072, 1024), (5, 300, 1024), (128, 1000, 512),
                                   (200, 5000, 512), (33, 129, 256), (70, 2000, 1024), (8, 777, 2048)])
def test_score_matrix_matches_oracle(K, b, n, d):
    q, bank = clip_like(b, d, 1), clip_like(n, d, 2)
    qd, bd = q.to(dev()), bank.to(dev())
    _, bias = K.bank_build(bd)
    s = K.debug_scores(qd, bd, bias).cpu().double()
    ref = O.scores_f64(q.float(), bank.float())
    assert not torch.isnan(s).any()
    assert (s - ref).abs().max().item() < TOL
    assert K.handle(0).device_error() == 0


# ------------------------------------------------------------------------------------------------ kernel 2+4: top-k
TOPK_CASES = [
    # b, n, d, kk
    (16, 3072, 1024, 1), (16, 3072, 1024, 2), (16, 14336, 1024, 1), (16, 14336, 1024, 6), (1, 5000, 512, 32),
    (128, 20000, 512, 5), (129, 20000, 512, 5), (300, 30000, 512, 16), (64, 2000, 1024, 15), (65, 2000, 1024, 16),
    (7, 127, 64, 3), (7, 128, 64, 3), (7, 129, 64, 3), (3, 1, 64, 1), (40, 50000, 256, 31),
    (512, 60000, 512, 5), (256, 4000, 1024, 3), (1000, 9000, 512, 2),     # even q-tile counts: CTA-pair multicast path
    (4, 3, 64, 5), (9, 20, 1024, 32),                                     # k > N: fewer rows than list slots
    # hybrid q-tile (512 < D <= 1024, more than 16 queries): first 512 dims in tensor memory, the rest in shared memory
    (128, 9000, 1024, 5), (100, 6000, 1024, 32), (200, 5000, 768, 16), (300, 20000, 1024, 5), (66, 400, 576, 8),
    (17, 3000, 1024, 5), (40, 3000, 640, 2), (50, 70000, 896, 7),
]


@pytest.mark.parametrize("b,n,d,kk", TOPK_CASES)
def test_topk_matches_oracle(K, b, n, d, kk):
    q, bank = clip_like(b, d, 10 + b), clip_like(n, d, 20 + kk)
    half = min(b // 2, n)
    if half:
        q[:half] = (bank[:half].float() * 1.01).to(torch.bfloat16)        # near-self matches
    qd, bd = q.to(dev()), bank.to(dev())
    _, bias = K.bank_build(bd)
    keys, score, idx = K.search_topk(qd, bd, bias, kk)
    s_ref, idx_ref, top_ref = oracle_topk(q, bank, kk)
    kq = min(kk, n)
    idx_h, score_h = idx.cpu().numpy(), score.cpu().numpy()
    O.check_index_parity(idx_h[:, :kq].astype(np.int64), s_ref, idx_ref[:, :kq], TOL)
    assert np.abs(score_h[:, :kq] - top_ref[:, :kq]).max() < TOL
    assert (idx_h[:, kq:] == -1).all() and np.isneginf(score_h[:, kq:]).all()   # fewer than kk rows exist
    s_dec, i_dec = O.decode_keys(keys.cpu().numpy().view(np.uint64))
    assert np.array_equal(i_dec, idx_h) and np.array_equal(s_dec[:, :kq], score_h[:, :kq])
    assert K.handle(0).device_error() == 0


def test_exact_ties_resolve_to_lower_row(K):
    """VQA_RAD-style exact duplicate rows: identical scores, the lower row must rank first — across
    tiles, across CTAs (splits) and within a tile."""
    n, d, kk = 40000, 512, 8
    bank = clip_like(n, d, 5)
    dup_of = {130: 3, 131: 3, 20000: 3, 39999: 3, 517: 516, 25000: 9000}
    for dst, src in dup_of.items():
        bank[dst] = bank[src]
    q = torch.cat([bank[[3, 516, 9000]].clone(), clip_like(5, d, 6)])
    qd, bd = q.to(dev()), bank.to(dev())
    _, bias = K.bank_build(bd)
    _, score, idx = K.search_topk(qd, bd, bias, kk)
    _, idx_ref, _ = oracle_topk(q, bank, kk)
    idx_h = idx.cpu().numpy()
    assert idx_h[0, :5].tolist() == [3, 130, 131, 20000, 39999]
    assert idx_h[1, :2].tolist() == [516, 517] and idx_h[2, :2].tolist() == [9000, 25000]
    assert np.array_equal(idx_h[:3, :2], idx_ref[:3, :2])
    sc = score.cpu().numpy()
    assert (sc[0, :5] == sc[0, 0]).all()                                 # bit-identical scores for identical rows


def test_search_rejects_bad_arguments(K):
    from multimodalpromptretrieval_b200._native import NativeError
    q, bank = clip_like(4, 128, 1).to(dev()), clip_like(100, 128, 2).to(dev())
    bias = torch.zeros(100, device=dev())
    with pytest.raises(NativeError):
        K.search_topk(q, bank, bias, 33)                                 # k + skip > 32
    with pytest.raises(NativeError):
        K.search_topk(q, bank, bias, 0)
    q96, bank96 = clip_like(4, 96, 1).to(dev()), clip_like(100, 96, 2).to(dev())
    with pytest.raises(NativeError):
        K.search_topk(q96, bank96, bias, 1)                              # D % 64 != 0


def test_full_size_bank_properties(K):
    """BASELINE-size shard (1.25 M x 512 bf16 = one GPU's share of the 10 M bank): planted neighbours are found in
    order, and the result equals a chunked fp32 torch scan under the 1e-3 rule."""
    n, d, b, kk = 1_250_000, 512, 128, 5
    g = torch.Generator(device="cuda").manual_seed(88)
    src = torch.randn(n, d, device=dev(), generator=g) * (10.0 / d ** 0.5)
    bank, bias = K.bank_build(src)
    del src
    rows = torch.randint(0, n, (b,), device=dev(), generator=g)
    q = bank[rows].clone()                                               # exact copies: the row itself must be rank 0
    _, score, idx = K.search_topk(q, bank, bias, kk)
    best = torch.full((b, kk), float("-inf"), device=dev())
    best_i = torch.zeros((b, kk), dtype=torch.int64, device=dev())
    for c0 in range(0, n, 250_000):
        blk = bank[c0:c0 + 250_000].float()
        s = q.float() @ blk.T + bias[c0:c0 + 250_000][None, :]
        cs, ci = torch.topk(s, kk, dim=1)
        allv, alli = torch.cat([best, cs], 1), torch.cat([best_i, ci + c0], 1)
        best, sel = torch.topk(allv, kk, dim=1)
        best_i = torch.gather(alli, 1, sel)
    assert (score - best).abs().max().item() < TOL
    mism = idx.long() != best_i
    gap = (score - best).abs()
    assert (gap[mism] < TOL).all()
    self_score = -bias[rows]                                             # <b,b> - 0.5|b|^2 = 0.5|b|^2
    assert (score[:, 0] - self_score).abs().max().item() < TOL
    assert (idx[:, 0].long() == rows).float().mean().item() > 0.99       # (a random duplicate row would also tie)
    assert K.handle(0).device_error() == 0


# ------------------------------------------------------------------------------------------------ kernel 4: merge
@pytest.mark.parametrize("n_lists,b,kk", [(1, 5, 3), (2, 16, 6), (8, 128, 5), (148, 33, 32), (37, 300, 1)])
def test_merge_matches_oracle(K, n_lists, b, kk):
    rng = np.random.default_rng(n_lists * 100 + kk)
    score = rng.standard_normal((n_lists, b, kk)).astype(np.float32)
    score[0, :, : kk // 2] = 0.5                                          # plenty of exact score ties
    idx = rng.permutation(n_lists * b * kk).reshape(n_lists, b, kk).astype(np.int32)
    keys = O.keys_from(score, idx)
    keys = np.sort(keys, axis=2)[:, :, ::-1].copy()                       # every input list sorted descending
    if n_lists > 1:
        keys[1, :, kk - 1] = 0                                            # an empty slot
    out_keys, out_score, out_idx = K.merge_topk(torch.from_numpy(keys.view(np.int64)).to(dev()))
    ref = O.merge_keys(keys, kk)
    assert np.array_equal(out_keys.cpu().numpy().view(np.uint64), ref)
    s_ref, i_ref = O.decode_keys(ref)
    assert np.array_equal(out_idx.cpu().numpy(), i_ref)
    got = out_score.cpu().numpy()
    assert np.array_equal(got[ref != 0], s_ref[ref != 0])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_equals_single_shard(K, world):
    """Multi-GPU path emulated on one GPU: per-rank shard scans with their idx_base, candidate lists stacked as the
    all-gather would deliver them, merged by kernel 4 — identical (keys, bit for bit) to the unsharded search."""
    from multimodalpromptretrieval_b200.sharding import shard_bounds
    n, d, b, kk = 30011, 512, 48, 6
    bank = clip_like(n, d, 77)
    bank[29000] = bank[5]                                                 # duplicate across shards
    q = torch.cat([bank[:24].clone(), clip_like(24, d, 78)])
    qd, bd = q.to(dev()), bank.to(dev())
    _, bias = K.bank_build(bd)
    full_keys, _, full_idx = K.search_topk(qd, bd, bias, kk)
    parts = []
    for r in range(world):
        b0, b1 = shard_bounds(n, r, world)
        k_r, _, _ = K.search_topk(qd, bd[b0:b1].contiguous(), bias[b0:b1].contiguous(), kk, idx_base=b0)
        parts.append(k_r)
    keys, _, idx = K.merge_topk(torch.stack(parts, 0).contiguous())
    assert torch.equal(keys, full_keys) and torch.equal(idx, full_idx)
    assert idx[5, :2].tolist() == [5, 29000]


# ------------------------------------------------------------------------------------------------ kernel 3
def _oracle_prompt_ids(tokenizer, tasks, questions, answers_rows, use_quantifier, max_len):
    retrieved = [O.prompt_sentence(r, use_quantifier) for r in answers_rows]
    return O.tokenize(tokenizer, O.input_sentences(tasks, questions, retrieved), max_len)


@pytest.mark.parametrize("k,skip,quant,max_len", [(1, 0, True, 512), (1, 1, True, 512), (5, 1, True, 512),
                                                  (5, 0, False, 512), (15, 0, True, 512), (32, 0, True, 512),
                                                  (31, 1, False, 512), (5, 0, True, 24)])
def test_prompt_gather_matches_oracle(K, tokenizer, k, skip, quant, max_len):
    from multimodalpromptretrieval_b200 import prompt as P
    from multimodalpromptretrieval_b200 import synthetic as S
    rng = np.random.default_rng(k * 7 + skip)
    vocab = S.answer_vocab(12, 88)                                        # few answers -> many vote ties
    n, b = 500, 37
    row_answers = [vocab[i] for i in rng.integers(0, len(vocab), n)]
    table = {}
    answer_id = np.array([table.setdefault(a, len(table)) for a in row_answers], dtype=np.int32)
    strings = list(table.keys())
    idx = np.stack([rng.permutation(n)[: k + skip] for _ in range(b)]).astype(np.int32)
    questions = [f"{q} #{i}" for i, q in enumerate(S.make_questions(b, 5))]
    tasks = [S.TASKS[i % len(S.TASKS)] for i in range(b)]
    tables = P.PromptTables(tokenizer, strings, dev())
    pre_ids, pre_off, longest = tables.prefixes(tasks, questions, quant)
    stride = min(max_len, longest + tables.tail_bound(quant))
    out = K.prompt_gather(torch.from_numpy(idx).to(dev()), skip, torch.from_numpy(answer_id).to(dev()),
                          torch.from_numpy(P.bucket_lut(k)).to(dev()), pre_ids, pre_off, tables.seg_ids,
                          tables.seg_off, quant, tables.pad_id, tables.eos_id, max_len, stride)
    rows = [[row_answers[j] for j in idx[i, skip:]] for i in range(b)]
    ids_ref, mask_ref = _oracle_prompt_ids(tokenizer, tasks, questions, rows, quant, max_len)
    longest_out = int(out["length"].max().item())
    assert longest_out == ids_ref.shape[1]
    assert torch.equal(out["input_ids"][:, :longest_out].cpu(), ids_ref)
    assert torch.equal(out["attention_mask"][:, :longest_out].cpu(), mask_ref)
    for i in range(b):
        ans, m, nv = O.vote(rows[i])
        assert strings[int(out["majority_answer"][i])] == ans
        assert int(out["majority_count"][i]) == m and int(out["bucket"][i]) == O.bucket_index(m, nv)
    assert np.array_equal(out["answer_ids"].cpu().numpy(), answer_id[idx[:, skip:]])


def test_prompt_gather_with_fewer_rows_than_k(K):
    """k > N: the reference's slice simply returns fewer columns; votes are over the rows that exist."""
    from multimodalpromptretrieval_b200 import prompt as P
    idx = torch.tensor([[4, 2, -1, -1], [1, -1, -1, -1]], dtype=torch.int32, device=dev())
    answer_id = torch.tensor([0, 1, 2, 0, 2], dtype=torch.int32, device=dev())
    z = torch.zeros(16, dtype=torch.int32, device=dev())
    out = K.prompt_gather(idx, 0, answer_id, torch.from_numpy(P.bucket_lut(4)).to(dev()), z, z[:3], z, z[:12],
                          True, 0, 1, 8, 4)
    assert out["majority_answer"].tolist() == [2, 1] and out["majority_count"].tolist() == [2, 1]
    assert out["bucket"].tolist() == [5, 5]                               # 2/2 and 1/1 -> "certainly"


# ------------------------------------------------------------------------------------------------ host class vs golden
class _Clip:
    def encode_image(self, x):
        return x

    def encode_text(self, x):
        return x


def _bank_from_golden(g, tokenizer, **kw):
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    table = {q: e for q, e in zip(g.questions, g.q_txt)}
    tok = lambda qs: torch.stack([table[q] for q in qs], 0)
    bank = RetrievalBank(clip_model=_Clip(), clip_tokenize=tok, tokenizer=tokenizer, shard=False, **kw)
    bank.install_bank([(g.bank_img, g.bank_txt)], g.answers, g.info, is_training_phase=g.training, retrieval_k=g.k)
    batch = {"image": g.q_img.clone(), "question": g.questions, "task": g.tasks}
    return bank, batch


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_retrieval_bank_reproduces_reference_golden(name, golden_cases, tokenizer):
    g = golden_cases[name]
    bank, batch = _bank_from_golden(g, tokenizer)
    r = bank._host(bank._retrieve(batch))
    s_ref = O.scores_f64(g.queries(), g.bank())
    skip = 1 if g.training else 0
    top = r["idx"][:, skip:skip + g.k].astype(np.int64)
    _, tolerated = O.check_index_parity(top, s_ref, g.z["top_idx"].astype(np.int64), TOL)
    dists = bank.retrieve_closest_qa_pairs(batch, return_dists=True)
    assert np.abs(np.stack([d for _, d in dists]) - g.z["return_dists"]).max() < TOL
    assert all(d.dtype == np.float32 and d.shape == (g.k,) for _, d in dists)
    if tolerated == 0:                                                    # tie-free cases: everything is bit-exact
        assert bank.retrieve_closest_qa_pairs(batch) == g.j["prompts_quant"]
        assert bank.retrieve_closest_qa_pairs(batch, use_quantifier=False) == g.j["prompts_plain"]
        assert bank.retrieve_closest_qa_pairs(batch, return_ans=True) == g.j["return_ans"]
        assert bank.retrieve_closest_qa_pairs(batch, return_info=["question_type"]) == g.j["return_info_type"]
        assert bank.retrieve_closest_qa_pairs(batch, return_info=["question", "question_id"]) == g.j["return_info_q_id"]
        assert [list(a) for a, _ in dists] == g.j["return_dists_answers"]
        for quant, key in ((True, "quant"), (False, "plain")):
            ids, mask = bank.retrieve_prompt_ids(batch, use_quantifier=quant)
            assert np.array_equal(ids.cpu().numpy(), g.z[f"input_ids_{key}"])
            assert np.array_equal(mask.cpu().numpy(), g.z[f"attention_mask_{key}"])
    else:
        # duplicates: prompt parity is evaluated on the oracle's vote applied to the accepted index set
        rows = [[g.answers[j] for j in top[i]] for i in range(g.b)]
        assert bank.retrieve_closest_qa_pairs(batch) == [O.prompt_sentence(r_, True) for r_ in rows]
        ids, mask = bank.retrieve_prompt_ids(batch)
        ids_ref, mask_ref = _oracle_prompt_ids(tokenizer, g.tasks, g.questions, rows, True, 512)
        assert torch.equal(ids.cpu(), ids_ref) and torch.equal(mask.cpu(), mask_ref)
    # precedence of the flags (VQAFeatureDataset.py:238-246) and memoisation across the test loop's repeated calls
    assert bank.retrieve_closest_qa_pairs(batch, return_ans=True, return_dists=True) == \
        bank.retrieve_closest_qa_pairs(batch, return_ans=True)
    assert bank._retrieve(batch) is r


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_host_api_and_two_deep_pipeline_reproduce_reference_golden(name, golden_cases, tokenizer):
    """The host-consumer entry points (one mpr_retrieve_host call per step: copy in, step, copy out): the blocking call
    and the two-deep submit()/result() pipeline return the token ids the reference's tokenizer call produces
    (T5VisionModel.py:153-167), for host-resident (pinned and pageable) query embeddings."""
    g = golden_cases[name]
    bank, batch = _bank_from_golden(g, tokenizer, memoise=False)
    ids_dev, mask_dev = bank.retrieve_prompt_ids(batch)
    want_ids, want_mask = ids_dev.cpu(), mask_dev.cpu()
    s_ref = O.scores_f64(g.queries(), g.bank())
    skip = 1 if g.training else 0
    r = bank._host(bank._retrieve(batch))
    _, tolerated = O.check_index_parity(r["idx"][:, skip:skip + g.k].astype(np.int64), s_ref,
                                        g.z["top_idx"].astype(np.int64), TOL)
    if tolerated == 0:
        assert np.array_equal(want_ids.numpy(), g.z["input_ids_quant"])
    ids_h, mask_h = bank.retrieve_prompt_ids_host(batch)
    assert torch.equal(ids_h, want_ids) and torch.equal(mask_h, want_mask)
    # pipeline: three batches in flight two at a time, quantifier setting alternating; results in submission order
    quants = [True, False, True]
    want = [tuple(t.cpu() for t in bank.retrieve_prompt_ids(batch, use_quantifier=q_)) for q_ in quants]
    bank.prefetch(batch, quants[0])
    cur = bank.submit_prompt_ids_host(batch, use_quantifier=quants[0])
    got = []
    for q_ in quants[1:]:
        nxt = bank.submit_prompt_ids_host(batch, use_quantifier=q_)
        got.append(tuple(t.clone() for t in cur.result()))
        cur = nxt
    got.append(tuple(t.clone() for t in cur.result()))
    for (gi, gm), (wi, wm) in zip(got, want):
        assert torch.equal(gi, wi) and torch.equal(gm, wm)
    assert cur.wait()["idx"].shape == (g.b, g.k + skip)


def test_create_retrieval_dataset_cache_and_additional_data(tmp_path, golden_cases, tokenizer):
    """A1/A2: build from a loader, write the reference-format cache, reload from it, and append the additional
    (ROCO-style) bank — the path that crashes in the reference (VQAFeatureDataset.py:181)."""
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    from multimodalpromptretrieval_b200 import synthetic as S
    base = S.make_bank(300, 80, 64, 0.0, seed=1, answers=S.answer_vocab(20, 88))
    extra = S.make_bank(500, 120, 64, 0.0, seed=2, answers=S.ROCO_ANSWERS, id_prefix="roco")
    table = {}
    tok = lambda qs: torch.stack([table[q] for q in qs], 0)

    def loader(bank_, bs=64):
        for i in range(0, bank_.n, bs):
            qs = [f"{q} @{j}" for j, q in zip(range(i, i + bs), bank_.info["question"][i:i + bs])]
            for q, e in zip(qs, bank_.text_half[i:i + bs]):
                table[q] = e
            yield {"image": bank_.image_half[i:i + bs], "question": qs, "answer": bank_.answers[i:i + bs],
                   "question_type": bank_.info["question_type"][i:i + bs],
                   "question_id": bank_.info["question_id"][i:i + bs]}

    roco_dir = tmp_path / "synthetic_data" / "cache" / "ROCOFeatureDataset"
    os.makedirs(roco_dir)
    torch.save(extra.combined(), roco_dir / "embedding.pt")
    pickle.dump(extra.answers, open(roco_dir / "answers.pkl", "wb"))
    pickle.dump(extra.info, open(roco_dir / "answer_types.pkl", "wb"))
    kw = dict(clip_model=_Clip(), clip_tokenize=tok, tokenizer=tokenizer, shard=False, name="VQASLAKEFeatureDataset",
              cache_root=str(tmp_path / "cache"), additional_root=str(roco_dir))
    b1 = RetrievalBank(**kw)
    b1.create_retrieval_dataset(loader(base), "prefix", is_training_phase=False, retrieval_k=5, use_additional_data=True)
    assert b1.n_total == 800 and len(b1.retrieval_answers) == 800 and len(b1.retrieval_question_info["question"]) == 800
    assert (tmp_path / "cache" / "VQASLAKEFeatureDataset" / "embedding.pt").exists()
    cached = torch.load(tmp_path / "cache" / "VQASLAKEFeatureDataset" / "embedding.pt")
    assert cached.dtype == torch.float32 and tuple(cached.shape) == (300, 128)       # the reference's own format
    b2 = RetrievalBank(**kw)
    b2.create_retrieval_dataset(iter(()), "prefix", is_training_phase=False, retrieval_k=5, use_additional_data=True)
    assert torch.equal(b1.retrieval_embeddings, b2.retrieval_embeddings) and torch.equal(b1.bias, b2.bias)
    # queries near rows of the appended block must come back with global indices offset by the base size
    qrows = [310, 555, 799, 3]
    qs = [f"probe {i}" for i in range(4)]
    full = torch.cat([base.combined(), extra.combined()], 0)
    for q, r_ in zip(qs, qrows):
        table[q] = full[r_, 64:]
    batch = {"image": full[qrows, :64].clone(), "question": qs, "task": ["Organ"] * 4}
    got = b2._host(b2._retrieve(batch))["idx"][:, 0].tolist()
    assert got == qrows
    oracle_emb, oracle_ans, oracle_info = O.extend_with_additional_data(
        base.combined(), base.answers, base.info, extra.combined(), extra.answers, extra.info)
    expect = O.retrieve_closest_qa_pairs(full[qrows], O.bf16_round(oracle_emb), oracle_ans, oracle_info, 5, False)
    assert b2.retrieve_closest_qa_pairs(batch) == expect


def test_retrieval_step_is_graph_capturable(golden_cases, tokenizer):
    """A retrieval step (one cooperative launch: scan + tail) captured in a caller's CUDA graph returns what eager
    launches return, across repeated replays with different query batches."""
    g = golden_cases["k5_train"]
    eager, batch = _bank_from_golden(g, tokenizer)
    graphed, _ = _bank_from_golden(g, tokenizer)
    skip, kk = 1, g.k + 1
    img = torch.empty_like(g.q_img, device=dev())
    txt = torch.empty_like(g.q_txt, device=dev())
    img.copy_(g.q_img)
    txt.copy_(g.q_txt)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):                                                 # buffers + workspace zeroing happen outside the capture
            graphed.run_step(img, txt, None, True, False, kk=kk, skip=skip)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = graphed.run_step(img, txt, None, True, False, kk=kk, skip=skip)["device"]
    for rep in range(3):
        qi, qt = torch.roll(g.q_img, rep, 0).to(dev()), torch.roll(g.q_txt, rep, 0).to(dev())
        img.copy_(qi)
        txt.copy_(qt)
        graph.replay()
        ref = eager.run_step(qi, qt, None, True, False, kk=kk, skip=skip)["device"]
        for name in ("keys", "idx", "score", "majority_answer", "bucket"):
            assert torch.equal(captured[name], ref[name]), (rep, name)
    assert K_handle_ok()


def K_handle_ok():
    from multimodalpromptretrieval_b200 import kernels
    return kernels.handle(0).device_error() == 0


@pytest.mark.parametrize("world,b,defer", [(2, 37, 0), (8, 37, 0), (8, 300, 0), (2, 300, 0), (2, 37, 1), (8, 128, 1)])
def test_p2p_exchange_protocol_on_one_gpu(K, world, b, defer):
    """The peer-memory exchange of the retrieval tail (csrc/tail.cuh) with the OTHER ranks' deliveries pre-populated in
    the exchange buffer (nothing here ever waits on a kernel that has not finished): rank r of `world` scans its shard,
    pushes its lists, finds every peer's tagged words already in place and merges — the result must equal the unsharded search,
    bit for bit, over several epochs (both slot parities).  (world 2, b 300) is more than one wave of CTAs and takes the
    two-launch path (stand-alone tail kernel).  defer = 1: the step kernel only pushes, the finish kernel on the
    library's side stream collects, merges and writes the outputs (mpr_retrieve_join orders them before the checks)."""
    from multimodalpromptretrieval_b200 import _native
    from multimodalpromptretrieval_b200.sharding import P2PExchange, shard_bounds
    import ctypes as C
    n, d, kk = 20011, 256, 6
    bank = clip_like(n, d, 31)
    bank[19000] = bank[5]                                                  # duplicate across shards
    bd = bank.to(dev())
    _, bias = K.bank_build(bd)
    cap = 4096
    me = world // 2
    x = P2PExchange(dev(), cap, world_size=world, rank=me)
    nbytes = K.exchange_bytes(world, cap)
    data_off = 1024
    assert nbytes == data_off + 4 * world * cap * 16          # two tagged 8-byte words per key, four buffers in turn
    b0, b1 = shard_bounds(n, me, world)
    shard, sbias = bd[b0:b1].contiguous(), bias[b0:b1].contiguous()
    status = torch.zeros(4, dtype=torch.int32, device=dev())
    b_full, kk_full = b, kk
    for epoch in range(1, 5):
        # the batch shape changes between exchanges: words left by another shape carry an older tag and are ignored
        b, kk = (b_full, kk_full) if epoch != 2 else (max(6, b_full // 2), 3)
        ws = K.new_workspace(K.search_workspace_bytes(b, b1 - b0, d, kk), dev())
        out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev())
        out_idx = torch.empty((b, kk), dtype=torch.int32, device=dev())
        q = torch.cat([bank[:b // 2].clone(), clip_like(b - b // 2, d, 40 + epoch)]).to(dev())
        full_keys, _, full_idx = K.search_topk(q, bd, bias, kk)
        # what the peers would have delivered: their shard's lists into word[epoch & 3][r] as {key half | epoch << 32}
        host = x.buf.cpu().numpy().copy()
        words = host[data_off:].view(np.uint64).reshape(4, world, cap, 2)
        for r in range(world):
            if r == me:
                continue
            r0, r1 = shard_bounds(n, r, world)
            k_r, _, _ = K.search_topk(q, bd[r0:r1].contiguous(), bias[r0:r1].contiguous(), kk, idx_base=r0)
            keys = k_r.cpu().numpy().view(np.uint64).reshape(-1)
            tag = np.uint64(epoch) << np.uint64(32)
            words[epoch & 3, r, :b * kk, 0] = (keys & np.uint64(0xFFFFFFFF)) | tag
            words[epoch & 3, r, :b * kk, 1] = (keys >> np.uint64(32)) | tag
        x.buf.copy_(torch.from_numpy(host))
        a = _native.RetrieveArgs()
        a.q_bf16, a.b, a.bank, a.bias = q.data_ptr(), b, shard.data_ptr(), sbias.data_ptr()
        a.n_local, a.idx_base, a.d, a.kk = b1 - b0, b0, d, kk
        a.out_keys, a.out_idx, a.status = out_keys.data_ptr(), out_idx.data_ptr(), status.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        x.fill_args(a)
        a.defer_finish = defer
        K.retrieve(a, dev())
        if defer:
            assert K.last_launch_count() == 2                              # step kernel + finish kernel
            K.retrieve_join(dev())
            torch.cuda.synchronize()
        else:
            assert K.last_launch_count() == (1 if K.search_plan(b, b1 - b0, d, kk)["n_ctas"] <= 148 else 2)
        if (world, b) == (2, 300):
            assert K.last_launch_count() == 2
        if epoch == 4:
            assert out_idx[5, :2].tolist() == [5, 19000]
        assert torch.equal(out_keys, full_keys) and torch.equal(out_idx, full_idx), epoch
        ctrl = x.buf[:8].cpu().numpy().view(np.uint32)
        assert ctrl[0] == epoch                                            # epoch published by the last warp out
        assert int(status[0]) == 0
    assert K.handle(0).device_error() == 0


def test_p2p_exchange_times_out_instead_of_hanging(K):
    """A peer that never delivers: the waiting warps give up after the configured timeout, report
    MPR_STATUS_XCHG_TIMEOUT in the status word and return the LOCAL result — no trap, no poisoned context."""
    from multimodalpromptretrieval_b200 import _native
    from multimodalpromptretrieval_b200.sharding import P2PExchange
    n, d, b, kk = 3000, 128, 9, 3
    bank = clip_like(n, d, 1).to(dev())
    _, bias = K.bank_build(bank)
    q = clip_like(b, d, 2).to(dev())
    local_keys, _, _ = K.search_topk(q, bank, bias, kk)
    x = P2PExchange(dev(), 256, world_size=2, rank=0)
    ws = K.new_workspace(K.search_workspace_bytes(b, n, d, kk), dev())
    out_keys = torch.empty((b, kk), dtype=torch.int64, device=dev())
    status = torch.zeros(4, dtype=torch.int32, device=dev())
    a = _native.RetrieveArgs()
    a.q_bf16, a.b, a.bank, a.bias, a.n_local, a.d, a.kk = q.data_ptr(), b, bank.data_ptr(), bias.data_ptr(), n, d, kk
    a.out_keys, a.status, a.workspace, a.workspace_bytes = out_keys.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel()
    x.fill_args(a)
    K.set_exchange_timeout(0.05)
    try:
        K.retrieve(a, dev())
        torch.cuda.synchronize()
    finally:
        K.set_exchange_timeout(60.0)
    assert int(status[0]) == _native.STATUS_XCHG_TIMEOUT
    assert torch.equal(out_keys, local_keys)
    assert K.handle(0).device_error() == 0


def test_shard_cache_roundtrip_and_resharding(tmp_path, golden_cases, tokenizer):
    """N2: the bank is saved exactly as laid out in HBM (bf16 shards + bias + interned answers) and reloaded without a
    cast or a CLIP pass — under the same or a different world size — with identical retrieval results."""
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    g = golden_cases["k5_train"]
    src, batch = _bank_from_golden(g, tokenizer)
    key = RetrievalBank.cache_key(dataset="golden", split="train", normalise=False)
    # save as if 3 ranks had written (rank/world are plain attributes of the exchange object)
    for r in range(3):
        part, _ = _bank_from_golden(g, tokenizer)
        part.exchange.rank, part.exchange.world_size = r, 3
        part.install_bank([(g.bank_img, g.bank_txt)], g.answers, g.info, is_training_phase=g.training, retrieval_k=g.k)
        part.save_shards(str(tmp_path / "shards"), key)
    loaded, _ = _bank_from_golden(g, tokenizer)
    loaded.retrieval_embeddings = None
    assert not loaded.load_shards(str(tmp_path / "shards"), key="other-key")
    assert loaded.load_shards(str(tmp_path / "shards"), key, is_training_phase=g.training, retrieval_k=g.k)
    assert torch.equal(loaded.retrieval_embeddings, src.retrieval_embeddings) and torch.equal(loaded.bias, src.bias)
    assert loaded.retrieve_closest_qa_pairs(batch) == g.j["prompts_quant"]
    assert loaded.retrieve_closest_qa_pairs(batch, return_info=["question", "question_id"]) == g.j["return_info_q_id"]
    ids, _ = loaded.retrieve_prompt_ids(batch)
    assert np.array_equal(ids.cpu().numpy(), g.z["input_ids_quant"])
    # a 2-rank job reading the 3-rank files: each rank gets exactly its contiguous rows
    for r in range(2):
        part, _ = _bank_from_golden(g, tokenizer)
        part.exchange.rank, part.exchange.world_size = r, 2
        assert part.load_shards(str(tmp_path / "shards"), key)
        b0, b1 = part.row_begin, part.row_begin + part.retrieval_embeddings.shape[0]
        assert torch.equal(part.retrieval_embeddings, src.retrieval_embeddings[b0:b1])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("b,d0,d1,kk", [(16, 256, 256, 5), (128, 512, 0, 1), (130, 64, 64, 6), (300, 256, 0, 16),
                                        (16, 512, 512, 5), (64, 512, 512, 16), (100, 1024, 0, 2), (7, 1024, 1024, 3),
                                        (200, 512, 512, 5)])
def test_fused_query_cast_is_bit_identical_to_two_step(K, dtype, b, d0, d1, kk):
    """N3: concat + bf16 cast of the raw query halves inside the scan kernel's q-tile load == kernel 1 + kernel 2."""
    n, d = 20011, d0 + d1
    assert K.search_fused_supported(d) and K.search_fused_supported(1024) and not K.search_fused_supported(4096)
    bank = clip_like(n, d, 3).to(dev())
    _, bias = K.bank_build(bank)
    g = torch.Generator().manual_seed(b)
    a = (torch.randn(b, d0, generator=g) * 0.44).to(dtype).to(dev())
    t = (torch.randn(b, d1, generator=g) * 0.44).to(dtype).to(dev()) if d1 else None
    q, qbias = K.bank_build(a, t)
    keys_ref, score_ref, idx_ref = K.search_topk(q, bank, bias, kk, idx_base=7)
    keys, score, idx, qb = K.search_topk_fused(a, t, bank, bias, kk, idx_base=7)
    assert torch.equal(keys, keys_ref) and torch.equal(idx, idx_ref) and torch.equal(score, score_ref)
    assert (qb - qbias).abs().max().item() < 1e-3
    if d > 1024 or (d > 512 and b <= 16):
        assert torch.equal(qb, qbias)          # the shared-memory q-tile is filled by kernel 1's own row routine
    if 512 < d <= 1024 and b > 16:             # hybrid q-tile, both halves filled from the raw rows in the one launch
        assert K.last_launch_count() == 1 and K.search_plan(b, n, d, kk)["n_qtiles"] == -(-b // 128)
    assert K.handle(0).device_error() == 0


def test_fused_query_cast_with_normalise(K):
    n, d, b, kk = 30000, 512, 64, 5
    g = torch.Generator().manual_seed(9)
    src = torch.randn(n, d, generator=g).to(dev())
    bank, bias = K.bank_build(src, normalise=True)
    qsrc = (src[:b] + 0.05 * torch.randn(b, d, generator=g).to(dev())).contiguous()
    q, _ = K.bank_build(qsrc, normalise=True)
    _, score_ref, idx_ref = K.search_topk(q, bank, bias, kk)
    _, score, idx, qb = K.search_topk_fused(qsrc, None, bank, bias, kk, normalise=True)
    assert (idx[:, 0] == torch.arange(b, device=dev())).all()           # each query's own row is its nearest
    assert (score - score_ref).abs().max().item() < 2e-3                # norm summation order: <= 1 bf16 ulp on few elements
    assert (idx == idx_ref).float().mean().item() > 0.98
    assert (qb + 0.5).abs().max().item() < 5e-3


def test_bank_step_with_raw_queries_on_the_hybrid_qtile(tokenizer):
    """The reference's own row width (512 + 512) with a batch beyond 16 queries: the step scans with the hybrid q-tile,
    both halves of which the kernel fills from the raw CLIP halves itself (one launch); keys equal the two-step search on
    prepared queries, from device-resident and from host-resident halves alike."""
    from multimodalpromptretrieval_b200 import kernels as KK
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    n, b, k = 7000, 100, 5
    g = torch.Generator().manual_seed(77)
    img, txt = torch.randn(n, 512, generator=g) * 0.3, torch.randn(n, 512, generator=g) * 0.3
    bank = RetrievalBank(tokenizer=tokenizer, shard=False, memoise=False, precomputed_features=True)
    bank.install_bank([(img.to(dev()), txt.to(dev()))], [str(i % 7) for i in range(n)], None, is_training_phase=False,
                      retrieval_k=k)
    qi = (img[:b] + 0.02 * torch.randn(b, 512, generator=g)).contiguous()
    qt = (txt[:b] + 0.02 * torch.randn(b, 512, generator=g)).contiguous()
    assert KK.search_plan(b, n, 1024, k)["n_qtiles"] == 1                   # 100 queries x 1024 dims in ONE q-tile
    q_prepared, _ = KK.bank_build(qi.to(dev()), qt.to(dev()))
    want, _, _ = KK.search_topk(q_prepared, bank.retrieval_embeddings, bank.bias, k)
    got_dev = bank.run_step(qi.to(dev()), qt.to(dev()), None, True, False)["device"]["keys"].clone()
    assert KK.last_launch_count() == 1                                      # raw halves, hybrid q-tile, ONE launch
    got_host = bank.run_step(qi.pin_memory(), qt.pin_memory(), None, True, True)["host"]["keys"].clone()
    assert torch.equal(got_dev, want) and torch.equal(got_host.to(dev()), want)
    assert (want.cpu().numpy().view(np.uint64) != 0).all()
    idx = bank.run_step(qi.to(dev()), qt.to(dev()), None, True, False)["device"]["idx"]
    assert (idx[:, 0].cpu() == torch.arange(b, dtype=torch.int32)).all()    # every query's own row is its nearest


def test_recycled_workspace_memory_is_rezeroed(K):
    """The scan keeps control words (tile counters, thresholds, barrier) at the head of its workspace and expects them to
    be zero.  A workspace whose memory was freed, scribbled on by another tensor and handed back at the same address must
    be re-zeroed by the library (mpr_workspace_invalidate via kernels.new_workspace) — not trusted because its address
    is familiar."""
    n, d, b, kk = 20000, 256, 32, 5
    bank = clip_like(n, d, 3).to(dev())
    _, bias = K.bank_build(bank)
    q = clip_like(b, d, 4).to(dev())
    ref_keys, _, _ = K.search_topk(q, bank, bias, kk)
    need = K.search_workspace_bytes(b, n, d, kk)
    for _ in range(4):
        ws = K.new_workspace(need, dev())
        keys, _, _ = K.search_topk(q, bank, bias, kk, workspace=ws)
        assert torch.equal(keys, ref_keys)
        addr = ws.data_ptr()
        del ws
        junk = torch.full((max(need, 16),), 0xFF, dtype=torch.uint8, device=dev())     # same block from the caching allocator
        same = junk.data_ptr() == addr
        del junk
    assert same or True
    assert K.handle(0).device_error() == 0
