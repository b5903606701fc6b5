"""Pin the oracle (oracle/captioner.py) against the fixtures the UNMODIFIED reference produced
(tests/golden/*.npz, made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import captioner as oc
from oracle import ref_harness


def _oracle(g):
    spec, gpt, mapper, task, x = gu.rebuild(g)
    gu.check_fingerprint(g, gpt, mapper)
    return oc.CaptionOracle(spec, gpt, mapper, task_prefix_embeds=task), x


@pytest.mark.parametrize("name", ["tiny_mlp_eos", "tiny_tfm", "tiny_mlp_task"])
def test_tiny_cases_token_exact(name):
    g = gu.load(name)
    o, x = _oracle(g)
    N = int(g["max_length"])
    eos = int(g["eos"])
    assert eos == oc.EOS_TOKEN_ID
    for sl, L, ref_ids in gu.golden_batches(g):
        for kv in (False, True):  # the reference's quadratic loop, and the cached form used for big sets
            ids = o.generate(x[sl], N, kv_cache=kv)
            assert ids.shape == (sl.stop - sl.start, L), (name, sl, kv)
            assert np.array_equal(ids.numpy(), ref_ids), (name, sl, kv)
        ids_hf = o.generate(x[sl], N, backend="hf")
        assert np.array_equal(ids_hf.numpy(), ref_ids)
        assert oc.trim_length(torch.from_numpy(_full_rows(ref_ids, N, eos)), N, eos) == L
    # step-0 logits (prefill, no cascade) within fp32 reordering noise
    rows = g["logits0"].shape[0]
    _, logs = o.generate(x[:rows], 1, return_logits=True)
    np.testing.assert_allclose(logs[0].numpy(), g["logits0"], rtol=0, atol=2e-5)


def _full_rows(ids, N, eos):
    out = np.full((ids.shape[0], N), eos, np.int64)
    out[:, : ids.shape[1]] = ids
    return out


def test_eos_fixture_actually_exercises_early_stop():
    g = gu.load("tiny_mlp_eos")
    lens = g["batch_lens"]
    assert lens.min() < int(g["max_length"]) and lens.max() == int(g["max_length"])
    assert (g["ids"] == int(g["eos"])).any()


def test_trim_rule_properties():
    eos = oc.EOS_TOKEN_ID
    t = torch.tensor
    assert oc.trim_length(t([[1, 2, 3], [4, 5, 6]]), 3) == 3  # nobody finished
    assert oc.trim_length(t([[1, eos, eos], [eos, eos, eos]]), 3) == 2  # max first-EOS + 1
    assert oc.trim_length(t([[1, eos, eos], [4, 5, 6]]), 3) == 3  # one unfinished row keeps the loop alive
    assert oc.trim_length(t([[1, 2, eos]]), 3) == 3
    assert oc.trim_length(torch.empty(2, 0, dtype=torch.long), 0) == 0


@pytest.mark.skipif(not gu.have("c1_small_mlp_b64"), reason="fixture not generated")
def test_c1_small_mlp_tokens():
    """BASELINE.json config 1 (64 rows, 30 tokens).  Cached form on all rows; the reference's cache-less
    loop restated on the first 4 rows (the full 64 take ~40 s of CPU)."""
    g = gu.load("c1_small_mlp_b64")
    o, x = _oracle(g)
    ref_ids = g["ids"].astype(np.int64)
    ids = o.generate(x, 30, kv_cache=True)
    assert np.array_equal(ids.numpy(), ref_ids)
    ids4 = o.generate(x[:4], 30, kv_cache=False)
    assert np.array_equal(ids4.numpy(), ref_ids[:4])
    _, logs = o.generate(x[:2], 1, return_logits=True)
    np.testing.assert_allclose(logs[0].numpy(), g["logits0"], rtol=0, atol=2e-5)


@pytest.mark.skipif(not gu.have("c4_large_mlp"), reason="fixture not generated")
@pytest.mark.slow
def test_c4_large_tokens():
    g = gu.load("c4_large_mlp")
    o, x = _oracle(g)
    ids = o.generate(x, 30, kv_cache=True)
    assert np.array_equal(ids.numpy(), g["ids"].astype(np.int64))


@pytest.mark.skipif(not gu.have("c1_small_mlp_b64"), reason="fixture not generated")
def test_bf16_emulation_documents_the_random_init_margin_problem():
    """With RANDOM-INIT weights the greedy margins are small (median top-2 gap 0.24, min 3e-3): merely
    storing GEMM operands in bfloat16 (fp32 accumulation, emulated here on the CPU) changes a large share of 30-token
    captions while the logits stay within 1e-2.  This is why the bf16 mode's caption agreement with the fp32 reference is
    reported rather than asserted at 99 %, and why the BF16X2 mode exists (precision modes)."""
    g = gu.load("c1_small_mlp_b64")
    o, x = _oracle(g)
    ref = torch.from_numpy(g["ids"].astype(np.int64))
    emu, logs = o.generate(x, 30, kv_cache=True, emulate_bf16=True, return_logits=True)
    match = float((emu == ref).all(dim=1).float().mean())
    assert 0.5 < match < 0.99, match
    rel = (logs[0][:2] - torch.from_numpy(g["logits0"])).abs().max() / torch.from_numpy(g["logits0"]).abs().max()
    assert rel < 1e-2


def test_beam_fixture_reproduces():
    """Beam search is not in the reference; its oracle is HF GenerationMixin on the same pinned weights."""
    g = gu.load("tiny_mlp_beam5")
    spec = gu.spec_of(g)
    o = oc.CaptionOracle(spec, *gu._modules(spec, oc.EOS_TOKEN_ID, 1.0))
    x = oc.synthetic_embeddings(int(g["n_rows"]), spec.embed_dim, 1)
    out = o.generate_beam(x, int(g["max_length"]), int(g["num_beams"]))
    assert np.array_equal(out.numpy(), g["ids"].astype(np.int64))


def _rat_db(g):
    rng = np.random.default_rng(int(g["db_seed"]))
    n_img, D = int(g["n_img"]), 512
    img = rng.standard_normal((n_img, D)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    counts = rng.integers(0, 7, n_img)
    assert np.array_equal(counts, g["counts"])
    owner = np.repeat(np.arange(n_img), counts)
    cap = rng.standard_normal((len(owner), D)).astype(np.float32)
    starts = np.concatenate([[0], np.cumsum(counts)])
    return img, cap, starts


def test_retrieval_restatement_matches_reference_functions():
    g = gu.load("rat_retrieval")
    img, cap, starts = _rat_db(g)
    rows_of = lambda i: list(range(starts[i], starts[i + 1]))  # noqa: E731
    for (k, i) in [(10, 4), (20, 6), (5, 1)]:
        aug, rows = oc.retrieve_and_aggregate(img, cap, rows_of, g["q"], top_i=i, top_k=k)
        ret = np.zeros((rows.shape[0], k, 512), np.float32)
        ret[rows >= 0] = cap[rows[rows >= 0]]
        assert np.array_equal(ret, g[f"ret_k{k}_i{i}"])
        np.testing.assert_allclose(aug, g[f"aug_k{k}_i{i}"], rtol=0, atol=1e-6)
    # the self-match filter fired for the rows that are exact DB entries
    hits = g["hits_k10_i4"]
    for qi, dbi in enumerate([5, 17, 100, 101, 1999, 0]):
        assert dbi not in hits[qi]


def test_flat_ip_tie_break_lowest_index():
    db = np.zeros((6, 4), np.float32)
    db[[1, 3, 4], 0] = 1.0  # three identical rows
    db[5, 0] = 2.0
    s, i = oc.flat_ip_search(db, np.array([[1, 0, 0, 0]], np.float32), 8)
    assert i[0].tolist() == [5, 1, 3, 4, 0, 2, -1, -1]
    assert s[0, -1] == -np.inf


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_equals_live_reference_on_fresh_inputs():
    """Where the reference is importable (build container), compare on inputs that are NOT in the fixtures."""
    ref, _ = ref_harness.import_reference()
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4, seed=3)
    gpt, mapper = oc.build_modules(spec)
    torch.manual_seed(3)
    from transformers import GPT2Config, GPT2LMHeadModel
    rgpt = GPT2LMHeadModel(GPT2Config(**spec.dims))
    rmap = ref.MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
    model = ref.ImageCaptioningModel(mapping_network=rmap, tokenizer=ref_harness.StubTokenizer(), gpt=rgpt).eval()
    x = oc.synthetic_embeddings(6, 64, seed=11)
    want = model.generate(image_embeddings=x, max_length=9, temperature=0.0)
    got = oc.CaptionOracle(spec, gpt, mapper).generate(x, 9)
    assert torch.equal(want, got)


@pytest.mark.parametrize("full,ref_made,rows", [("c2_small_mlp_full5000", "c2_small_mlp_first1024", 1024), ("c3_medium_tfm_full32", "c3_medium_tfm", 8),
                                                ("c4_large_mlp_full16", "c4_large_mlp", 4), ("c3_medium_tfm_beam5_full32", "c3_medium_tfm_beam5", 4),
                                                ("c3_medium_tfm_full256", "c3_medium_tfm", 8), ("c4_large_mlp_full256", "c4_large_mlp", 4)])
def test_full_fixtures_extend_the_reference_ones(full, ref_made, rows):
    """The full-size fixtures (tests/golden/make_golden_full.py: oracle, KV-cached; beam through HF) agree token for token with the
    fixtures the UNMODIFIED reference produced on every row they share -- same pinned weights (fingerprint), same embeddings."""
    a, b = gu.load(full), gu.load(ref_made)
    assert str(a["spec_gpt"]) == str(b["spec_gpt"]) and int(a["spec_prefix_length"]) == int(b["spec_prefix_length"])
    if "fp_abs_sum" in a and "fp_abs_sum" in b:
        assert abs(float(a["fp_abs_sum"]) - float(b["fp_abs_sum"])) <= 1e-9 * abs(float(b["fp_abs_sum"]))
    ia, ib = a["ids"].astype(np.int64), b["ids"].astype(np.int64)
    L = min(ia.shape[1], ib.shape[1])
    assert ib.shape[0] == rows and np.array_equal(ia[:rows, :L], ib[:, :L])
    if "min_gap" in a:  # the audit data: positive, and small somewhere in 5 000 captions (random-init margins)
        assert (a["min_gap"] > 0).all() and a["min_gap"].shape[0] == ia.shape[0]


def test_c5_fixture_is_consistent_with_the_restated_search_on_a_subset():
    """The config-5 fixture (float64 scores, (score desc, index asc)) against oc.flat_ip_search -- the restatement of IndexFlatIP the
    retrieval parity tests use -- on the first rows of the seeded database (the full 591 753-row check is the -m gpu test)."""
    g = gu.load("c5_retrieval_full")
    gen = torch.Generator().manual_seed(int(g["db_seed"]))
    img = torch.randn(int(g["n_img"]), 512, generator=gen)
    img /= img.norm(dim=-1, keepdim=True)
    q = oc.synthetic_embeddings(1024, 512, 1)
    q[:4] = img[torch.from_numpy(g["exact_rows"].astype(np.int64))]
    s, i = oc.flat_ip_search(img.numpy(), q[:16].numpy(), 15)
    ok = (i == g["img_idx"][:16].astype(np.int64)).all(axis=1) | (g["img_min_gap"][:16] < 1e-6)
    assert ok.all()
    np.testing.assert_allclose(s, g["img_scores"][:16], atol=2e-6, rtol=0)
    assert (g["img_scores"][:4, 0] > 0.9999).all()  # the exact database rows are their own best hit ...
    assert (g["rat_rows"][:4, 0] // 5 != g["exact_rows"]).all()  # ... and are filtered (faiss_store.py:160-163)


def test_engine_flips_on_c2_are_the_rows_the_roundings_predict():
    """Cross-check of the bf16x2 engine against a CPU emulation of ITS roundings (split hi + lo GEMM operands, fp16 q / k / v stores,
    exact LM head; tests/precision_screen.py scheme hx_bf16x2_kv16): on the first 1024 rows of configs[1] the emulation differs from
    the fp32 oracle on four captions, and the engine's committed parity report (profiles/r2ac_parity_report.jsonl, written by
    tests/test_gpu_fullsize.py::test_c2_all_5000_rows_through_the_job_api on a B200) lists exactly those four rows plus one more
    near-tie among its mismatches below row 1024 -- the engine's flips are the roundings', not a kernel's."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import precision_screen as ps

    report = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2ac_parity_report.jsonl")
    engine = None
    for line in open(report):
        r = json.loads(line)
        if r.get("test") == "c2_full5000" and r.get("dtype") == "bf16x2":
            engine = r
    assert engine is not None and engine["mismatched_rows"] <= 50
    engine_rows = {a["row"] for a in engine["audits"] if a["row"] < 1024}  # (the report keeps the first 12 mismatches in row order: complete below 1024)
    assert max(a["row"] for a in engine["audits"]) >= 1024
    o = oc.CaptionOracle(oc.ModelSpec())
    x = oc.synthetic_embeddings(5000)[:1024]
    ref = torch.from_numpy(gu.load("c2_small_mlp_full5000")["ids"][:1024].astype(np.int64))
    ra, rw, rkv, _, _, head = ps.SCHEMES["hx_bf16x2_kv16"]
    ids = ps.generate(o, x, 30, ra, rw, rkv, head)
    emulated_rows = set(torch.nonzero((ids != ref).any(dim=1)).flatten().tolist())
    assert 1 <= len(emulated_rows) <= 10
    assert emulated_rows <= engine_rows, (sorted(emulated_rows), sorted(engine_rows))
    assert len(engine_rows - emulated_rows) <= 2
