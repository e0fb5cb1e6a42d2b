"""CPU tests of the host-side logic and of the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from multimodalpromptretrieval_b200 import _native, prompt, sharding
from oracle import retrieval_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bucket_lut_matches_reference_formula_for_every_m_k():
    for k in range(1, 33):
        lut = prompt.bucket_lut(k).reshape(k + 1, k + 1)
        for n in range(1, k + 1):
            for m in range(1, n + 1):
                assert lut[n, m] == O.bucket_index(m, n), (k, n, m)
    assert prompt.bucket_lut(5).reshape(6, 6)[5, 3] == 3          # 0.6*5 = 3.0000000000000004 -> 3
    assert prompt.bucket_lut(1)[3] == 5                           # k = 1 is always "certainly" 


def test_prompt_string_matches_oracle():
    for quant in (True, False):
        for row in (["yes"], ["a", "b", "b"], ["left lung", "x", "left lung", "x", "y"]):
            ans, m, n = O.vote(row)
            assert prompt.prompt_string(O.bucket_index(m, n), ans, quant) == O.prompt_sentence(row, quant)


def test_tokenisation_by_concatenation_over_whole_vocabulary(tokenizer):
    """H4: tokens(prefix + retrieved sentence) == tokens(prefix+"I") | tokens(tail) | tokens(bucket) | tokens(answer)."""
    from multimodalpromptretrieval_b200 import synthetic as S
    answers = S.answer_vocab(500, 88) + S.ROCO_ANSWERS + ["", "  padded   answer ", "Mixed Case: x?"]
    segs = prompt.segment_strings(answers)
    seg_tok = tokenizer(segs, add_special_tokens=False)["input_ids"]
    questions = S.make_questions(40, 3) + ["", "ends with space ", "what?"]
    tasks = [S.TASKS[i % len(S.TASKS)] for i in range(len(questions))]
    for quant in (True, False):
        pre = tokenizer(prompt.prefix_texts(tasks, questions, quant), add_special_tokens=False)["input_ids"]
        sents, expect = [], []
        for qi in range(len(questions)):
            for ai in range(qi, len(answers), len(questions)):
                b = (qi + ai) % 6
                retrieved = prompt.prompt_string(b, answers[ai], quant)
                sents.append(f"Answer the {tasks[qi]} question: " + questions[qi] + retrieved)
                tail = seg_tok[prompt.SEG_QUANT] + seg_tok[prompt.SEG_BUCKET0 + b] if quant else seg_tok[prompt.SEG_PLAIN]
                expect.append(pre[qi] + tail + seg_tok[prompt.SEG_ANSWER0 + ai] + [tokenizer.eos_token_id])
        got = tokenizer(sents, add_special_tokens=True)["input_ids"]
        assert got == expect


def test_shard_bounds_partition_rows():
    for n in (0, 1, 7, 128, 1000, 10_000_000):
        for w in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= e - b <= -(-n // w) for b, e in spans)
            if n:
                for row in (0, n // 2, n - 1):
                    r = sharding.owner_of_row(row, n, w)
                    assert spans[r][0] <= row < spans[r][1]
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_key_encoding_orders_by_score_then_lower_row():
    score = np.array([[1.5, 1.5, -2.0, 0.0, -0.0, 3.25, -np.inf]], dtype=np.float32)
    idx = np.array([[7, 3, 1, 2, 9, 5, 4]], dtype=np.int32)
    keys = O.keys_from(score, idx)
    order = np.argsort(-keys.astype(np.float64), kind="stable")  # only for a rough look; exact check below
    srt = sorted(range(7), key=lambda i: (-score[0, i], idx[0, i]))
    exact = sorted(range(7), key=lambda i: int(keys[0, i]), reverse=True)
    # +0.0 and -0.0 differ in the ordered image (-0.0 < +0.0); everything else follows (score desc, row asc)
    assert [i for i in exact if i not in (3, 4)] == [i for i in srt if i not in (3, 4)]
    s2, i2 = O.decode_keys(keys)
    assert np.array_equal(s2.view(np.uint32), score.view(np.uint32)) and np.array_equal(i2, idx)
    merged = O.merge_keys(np.stack([keys[:, :4], keys[:, 3:7]], 0), 3)
    assert [int(x) for x in O.decode_keys(merged)[1][0]] == [5, 3, 7]
    del order


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodalpromptretrieval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "retrieval_oracle" not in src, f


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mpr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mpr_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_native.EXPORTS)
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mpr_abi_version.restype = ctypes.c_int
    assert lib.mpr_abi_version() == _native.ABI_VERSION == 2


def test_no_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodalpromptretrieval_b200 import kernels
    from multimodalpromptretrieval_b200.bank import RetrievalBank
    with pytest.raises(_native.NativeError):
        kernels.handle()
    with pytest.raises(RuntimeError):
        RetrievalBank()
    # mpr_create itself refuses: no device, no handle
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.mpr_create(0, ctypes.byref(h)) != 0 and not h.value


def test_fast_prefix_tokenisation_equals_hf_tokenizer(tokenizer):
    """prefix = cached tokens("Answer the {task} question:") + tokens(question + "I"|"The") through the direct
    sentencepiece batch call == the HF tokenizer on the whole prefix string (incl. empty / space-padded questions)."""
    from multimodalpromptretrieval_b200 import synthetic as S
    tables = prompt.PromptTables(tokenizer, ["yes", "no", "left lung"], torch.device("cpu"))
    questions = [f"{q} #{i}" for i, q in enumerate(S.make_questions(64, 11))] + ["", " lead", "trail ", "a  b"]
    tasks = [S.TASKS[i % len(S.TASKS)] for i in range(len(questions))]
    for quant in (True, False):
        ids, off = tables.prefix_tokens(tasks, questions, quant)
        ref = tokenizer(prompt.prefix_texts(tasks, questions, quant), add_special_tokens=False)["input_ids"]
        assert [ids[off[i]:off[i + 1]].tolist() for i in range(len(questions))] == ref
    assert tables.tail_bound(True) > tables.tail_bound(False) - 8 and tables.max_answer_len >= 1


def test_plan_shard_reads_reshards_across_world_sizes():
    n = 1003
    for w_saved in (1, 2, 3, 8):
        files = [sharding.shard_bounds(n, r, w_saved) for r in range(w_saved)]
        for w_load in (1, 2, 4, 5):
            seen = []
            for r in range(w_load):
                b, e = sharding.shard_bounds(n, r, w_load)
                plan = sharding.plan_shard_reads(files, b, e)
                pos = 0
                for i, first, cnt, dst in plan:
                    assert dst == pos and files[i][0] + first == b + dst and first + cnt <= files[i][1] - files[i][0]
                    pos += cnt
                    seen.extend(range(files[i][0] + first, files[i][0] + first + cnt))
                assert pos == e - b
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.plan_shard_reads([(0, 10)], 5, 20)


def test_header_is_plain_c():
    """The drop-in boundary must be bindable from C (cgo / JNI / ctypes all consume plain C declarations)."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "%s"\nint main(void) { mpr_handle_t h = 0; (void)h; return MPR_OK; }\n'
                % os.path.join(ROOT, "include", "mpr_b200.h"))
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f.name],
                         capture_output=True, text=True)
    os.unlink(f.name)
    assert res.returncode == 0, res.stderr


def test_shared_threshold_bound_never_drops_a_top_k_row():
    """The exactness argument behind the scan's shared admission thresholds (csrc/scan_topk.cuh), replayed on the host:
    lists cover disjoint rows; list (split, group) feeds word (slot * R + replica) with slot = (split + group * ns/2) % ns,
    replica = (split // ns) % R; a slot's value is the maximum over its replicas; the bound is the minimum over the ns
    slots.  Whatever subset of the lists has published so far, at least ns >= k+s distinct rows score >= the bound, so
    the (k+s)-th best score is >= it and dropping rows STRICTLY below it keeps every top-(k+s) row, ties included."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        kk = int(rng.integers(1, 33))
        ns = (kk + 3) & ~3
        rep = 1
        while rep < 4 and ns * rep * 2 <= (16 if kk <= 8 else 32):
            rep *= 2
        n_splits = int(rng.integers(ns, 149))
        groups = int(rng.integers(1, 3))
        rows_per_list = int(rng.integers(1, 40))
        n_lists = n_splits * groups
        # scores with many exact ties (few distinct values) in some trials
        levels = int(rng.choice([3, 17, 10_000]))
        scores = rng.integers(0, levels, size=(n_lists, rows_per_list)).astype(np.float64)
        published = rng.random(n_lists) < rng.choice([0.3, 0.7, 1.0])            # lists that have published so far
        words = np.full(ns * rep, -np.inf)
        for l_ in range(n_lists):
            if not published[l_]:
                continue
            split, grp = divmod(l_, groups)
            slot = (split + grp * (ns // 2)) % ns
            replica = (split // ns) % rep
            # a list publishes the best score among the rows it has seen (any prefix of its rows is a valid state)
            seen = int(rng.integers(1, rows_per_list + 1))
            words[slot * rep + replica] = max(words[slot * rep + replica], scores[l_, :seen].max())
        slot_val = words.reshape(ns, rep).max(axis=1)
        if np.isinf(slot_val).any():
            continue                                                             # some slot still empty: no bound yet
        bound = slot_val.min()
        flat = np.sort(scores.ravel())[::-1]
        if flat.size < kk:
            continue
        kth = flat[kk - 1]
        assert bound <= kth, (trial, kk, ns, rep)
        assert (scores >= bound).sum() >= kk                                     # everything that can matter survives
        assert not ((scores < bound) & (scores >= kth)).any()


PLAN_FIELDS = ("n_ctas", "n_splits", "n_qtiles", "n_stages", "smem_bytes", "q_tile", "sub_per_stage", "n_epi_groups", "q_tmem",
               "hybrid", "reg_list", "cand_cap", "q_box_rows", "workspace_bytes", "ns", "replicas")


def _plan(b, n, d, kk, sms=148):
    lib = _native.load()
    out = (ctypes.c_int32 * 16)()
    rc = lib.mpr_plan_host(sms, b, n, d, kk, out)
    return rc, dict(zip(PLAN_FIELDS, list(out)))


def test_launch_planner_invariants_over_the_shape_space():
    """The launch planner (csrc/mpr_abi.cu:make_plan) without a device: for every batch size / row width / list length the
    reference or BASELINE.json can produce, the plan fits shared memory, keeps a usable ring, covers the batch, and picks
    the q-tile placement the design describes."""
    for d in (64, 128, 256, 512, 576, 640, 768, 1024, 1536, 2048):
        for b in (1, 5, 16, 17, 33, 64, 65, 96, 128, 129, 200, 256, 300, 512, 1000, 4096):
            for kk in (1, 2, 5, 6, 8, 9, 15, 16, 31, 32):
                for n in (3, 3072, 1_250_000):
                    rc, p = _plan(b, n, d, kk)
                    assert rc == 0, (b, n, d, kk)
                    assert p["smem_bytes"] <= 232448, (b, n, d, kk, p)
                    assert p["n_stages"] >= 2 and p["sub_per_stage"] in (1, 2, 4)
                    assert p["n_qtiles"] * p["q_tile"] >= b and (p["n_qtiles"] - 1) * p["q_tile"] < b
                    assert 1 <= p["q_tile"] <= 128 and 1 <= p["n_splits"] and p["n_ctas"] == p["n_splits"] * p["n_qtiles"]
                    assert p["n_splits"] <= max(1, -(-n // 128))                       # no CTA without a tile of its own
                    assert p["reg_list"] == (1 if kk <= 8 else 0)
                    assert p["ns"] >= kk and p["ns"] % 4 == 0 and p["ns"] * p["replicas"] <= (16 if kk <= 8 else 32)
                    assert p["q_tmem"] == (1 if d <= 512 or p["hybrid"] else 0)
                    if p["hybrid"]:
                        assert 512 < d <= 1024 and b > 16 and p["n_stages"] >= 3
                        assert p["q_box_rows"] % 16 == 0 and p["q_box_rows"] >= min(b, 128) or p["q_box_rows"] == 128
                        assert (d // 64 - 8) * p["q_box_rows"] * 128 >= 32768           # room for the fill's scratch
                    if d <= 512:
                        assert p["q_box_rows"] == 0
                    if 640 <= d <= 1024 and b > 16:
                        assert p["hybrid"] == 1, (b, d, kk, p)
                    if p["n_qtiles"] == 1 and n >= 128 * 148:
                        assert p["n_ctas"] == 148                                        # one CTA per SM
    # the headline shapes
    rc, p = _plan(128, 10_000_000, 512, 5)
    assert (p["n_ctas"], p["n_qtiles"], p["q_tmem"], p["reg_list"], p["replicas"]) == (148, 1, 1, 1, 2) and p["n_stages"] * p["sub_per_stage"] >= 8
    rc, p = _plan(16, 1_062_912, 1024, 5)
    assert (p["n_qtiles"], p["q_tmem"], p["hybrid"]) == (1, 0, 0)
    rc, p = _plan(128, 1_000_000, 1024, 16)
    assert (p["n_qtiles"], p["hybrid"], p["q_box_rows"], p["n_epi_groups"]) == (1, 1, 128, 1)
    rc, p = _plan(4096, 1_048_576, 512, 5)
    assert p["n_qtiles"] == 32 and p["n_ctas"] % 148 == 0
    for bad in ((0, 100, 512, 5), (4, 100, 100, 5), (4, 100, 512, 33), (4, 100, 8192, 5)):
        assert _plan(*bad)[0] != 0


def test_ctypes_mirrors_match_the_header_structs(tmp_path):
    """``_native.RetrieveArgs`` / ``HostIO`` are written by hand next to ``mpr_retrieve_args`` / ``mpr_host_io``: size and
    every field offset must agree with what a C compiler makes of include/mpr_b200.h."""
    import shutil
    import subprocess
    from multimodalpromptretrieval_b200 import _native
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    pairs = [("mpr_retrieve_args", _native.RetrieveArgs), ("mpr_host_io", _native.HostIO)]
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "%s"' % os.path.join(ROOT, "include", "mpr_b200.h"),
             "int main(void) {"]
    for cname, cls in pairs:
        lines.append('printf("%%zu\\n", sizeof(%s));' % cname)
        for fname, _ in cls._fields_:
            lines.append('printf("%%zu\\n", offsetof(%s, %s));' % (cname, fname))
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    res = subprocess.run([gcc, "-std=c99", "-o", str(exe), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = []
    for _, cls in pairs:
        want.append(ctypes.sizeof(cls))
        want.extend(getattr(cls, fname).offset for fname, _ in cls._fields_)
    assert got == want


def test_word_cached_tokenisation_equals_direct_tokenisation(tokenizer):
    """encode_by_words (per-chunk cache) == the tokenizer on the whole string, cold and warm, incl. odd whitespace,
    punctuation-only chunks and non-ASCII text (which must take the direct path)."""
    from multimodalpromptretrieval_b200 import synthetic as S
    tables = prompt.PromptTables(tokenizer, ["yes"], torch.device("cpu"))
    texts = [f"{q} #{i}-{i * 7}I" for i, q in enumerate(S.make_questions(300, 21))]
    texts += ["", "I", " lead spaceThe", "trail I", "a\\tb cI", "x\\ny", "A  B   C", "a--b ?! ::", "1/2 ½ thingsI",
              "é accentI", "nb\\xa0spI", "ｆｕｌｌwidth textI", "  ", "what is   the  organ  ?I"]
    ref = tokenizer(texts, add_special_tokens=False)["input_ids"]
    assert tables.encode_by_words(texts) == ref                  # cold cache
    assert tables.encode_by_words(texts) == ref                  # warm cache
    assert tables.encode_by_words(list(reversed(texts))) == list(reversed(ref))
    assert all(w.isascii() and " " not in w for w in tables._word_cache)


def test_native_token_cache_assembly_equals_interpreter_path_and_hf(tokenizer):
    """csrc/token_cache.cpp (mpr_token_cache_assemble / _put — host code, no GPU): the prefix CSR assembled natively from
    cached chunk tokenisations == the interpreter's word-cache path == the HF tokenizer on the whole prefix string, for
    new and repeated questions, empty / space-padded text, both heads; non-ASCII text takes the interpreter path."""
    import threading
    from multimodalpromptretrieval_b200 import prompt as P
    from multimodalpromptretrieval_b200 import synthetic as S

    def tables(native):
        t = P.PromptTables.__new__(P.PromptTables)
        t.tokenizer, t.pad_id, t.eos_id = tokenizer, 0, 1
        t.encode, t._sp = P._fast_encoder(tokenizer)
        t._task_head, t._word_cache, t._starts_word = {}, {}, None
        t._native_cache, t._task_index, t._head_table = None, {}, None
        t._native_lock, t.use_native_cache = threading.Lock(), native
        return t

    a, b = tables(True), tables(False)
    tasks = [S.TASKS[i % len(S.TASKS)] for i in range(64)]
    for rep in range(3):
        qs = [f"{q} #{rep % 2}-{i}" for i, q in enumerate(S.make_questions(64, 7 + rep % 2))]
        qs[3], qs[4], qs[5], qs[63] = "  two  spaces here ", "", "x", " trailing "
        for quant in (True, False):
            ia, oa = a.prefix_tokens(tasks, qs, quant)
            ib, ob = b.prefix_tokens(tasks, qs, quant)
            assert np.array_equal(ia, ib) and np.array_equal(oa, ob)
            head = "I" if quant else "The"
            hf = tokenizer([f"Answer the {t} question: " + q + head for t, q in zip(tasks, qs)],
                           add_special_tokens=False)["input_ids"]
            assert [ia[oa[i]:oa[i + 1]].tolist() for i in range(64)] == hf
    assert a._native_cache is not None and _native.load().mpr_token_cache_size(a._native_cache) > 50
    qs[7] = "où est la lésion ?"
    ia, oa = a.prefix_tokens(tasks, qs, True)
    ib, ob = b.prefix_tokens(tasks, qs, True)
    assert np.array_equal(ia, ib) and np.array_equal(oa, ob)
    assert a.prefix_tokens([], [], True)[1].tolist() == [0]
