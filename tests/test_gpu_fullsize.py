"""Full-size parity of BASELINE.json configs 2-5 in EVERY arithmetic mode the benchmarks quote (VERDICT r1 item 1, row g), and
the reference's own objects through the boundary (item 6).

Fixtures: tests/golden/*_full*.npz, made by tests/golden/make_golden_full.py with the oracle's KV-cached restatement (pinned against
the reference-made fixtures by tests/test_oracle.py) and, for beam search, HF GenerationMixin.

north_star tolerances, asserted here:
  fp32   : token ids exact; a mismatching row must be a near-tie of the REFERENCE's own logits (its smallest top-2 gap over the
           caption, stored in the fixture, below 1e-4).
  bf16x2 : the "bf16 mode" contract -- at least 99 % of greedy captions identical to the fp32 reference (the mode the headline is quoted
           in; fp16 KV cache, operands hi + lo).
  bf16   : single-MMA operands.  Reported and asserted only against its CPU-emulated rounding model (>= 70 %): no single-MMA 16-bit
           scheme reaches 99 % on random-init weights (profiles/r2_precision_screen.jsonl).
"""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import gpu_util
from oracle import captioner as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def _report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def _audit_rows(g, got, ref):
    """Rows that differ, with the reference's own smallest top-2 logit gap over the caption and the first differing step."""
    bad = np.nonzero(~(got == ref).all(axis=1))[0]
    return [{"row": int(b), "first_diff_step": int(np.nonzero(got[b] != ref[b])[0][0]), "ref_min_gap": float(g["min_gap"][b]),
             "ref_min_gap_step": int(g["min_gap_step"][b])} for b in bad]


# floor per mode: fraction of captions identical to the fp32 reference
MODES = [("fp32", 0.999), ("bf16x2", 0.99), ("bf16", 0.70)]


@pytest.mark.parametrize("dtype,min_match", MODES)
def test_c2_all_5000_rows_through_the_job_api(dtype, min_match):
    """configs[1] at its stated size: 5 000 embeddings (val2017 size) from host memory through `generate_for_embeddings` --
    generate() calls of 1024 x 4 + a ragged 904, two batches in flight -- against the fixture's 5 000 x 30 tokens."""
    from gpt2_image_captioning_b200 import generate_for_embeddings
    g = gu.load("c2_small_mlp_full5000")
    model, _, x = gpu_util.product_model(g, dtype)
    assert x.shape == (5000, 512)
    ids = generate_for_embeddings(model, x, batch_size=1024, max_length=30, device=DEV, in_flight=2).numpy()
    ref = g["ids"].astype(np.int64)
    assert ids.shape == ref.shape == (5000, 30)
    row_ok = (ids == ref).all(axis=1)
    audits = _audit_rows(g, ids, ref)
    ragged_ok = float(row_ok[4096:].mean())
    _report(test="c2_full5000", dtype=dtype, rows=5000, caption_match=float(row_ok.mean()), token_match=float((ids == ref).mean()),
            ragged_batch_rows=904, ragged_batch_caption_match=ragged_ok, mismatched_rows=int((~row_ok).sum()), audits=audits[:12])
    assert row_ok.mean() >= min_match, f"{dtype}: {row_ok.mean():.4f} of 5 000 captions match the reference (need {min_match})"
    assert ragged_ok >= min_match - 0.01, f"{dtype}: ragged last batch only {ragged_ok:.4f}"
    if dtype == "fp32":
        for a in audits:
            assert a["ref_min_gap"] < 1e-4, f"fp32 mismatch that is not a near-tie of the reference: {a}"
    if dtype == "bf16x2":  # flips of the split mode sit at the reference's small margins, not anywhere
        gaps = np.array([a["ref_min_gap"] for a in audits] or [0.0])
        assert np.median(gaps) < 0.02, f"bf16x2 mismatches at large reference margins: {audits[:5]}"


@pytest.mark.parametrize("name,rows", [("c3_medium_tfm_full32", 32), ("c4_large_mlp_full16", 16)])
@pytest.mark.parametrize("dtype,min_match", MODES)
def test_c3_c4_greedy_full_fixture(name, rows, dtype, min_match):
    """configs[2] greedy (GPT-2 medium + 8-layer transformer mapper, P = 40, 32 rows) and configs[3] (GPT-2 large, E = 1024, 16 rows); the
    256-row sets of the same configs are in tests/test_zz_gpu_rates.py."""
    g = gu.load(name)
    model, _, x = gpu_util.product_model(g, dtype)
    ids = model.generate(image_embeddings=x.to(DEV), max_length=30, temperature=0.0).cpu().numpy()
    ref = g["ids"].astype(np.int64)
    assert ids.shape == ref.shape == (rows, 30)
    row_ok = (ids == ref).all(axis=1)
    audits = _audit_rows(g, ids, ref)
    _report(test="c3_c4_full", case=name, dtype=dtype, rows=rows, caption_match=float(row_ok.mean()), audits=audits[:8])
    if dtype == "bf16":
        assert row_ok.mean() >= 0.4, f"{name}/bf16: {row_ok.mean():.3f}"  # (16 - 32 rows: the 5 000-row test carries the statistic)
        return
    for a in audits:  # fp32 / bf16x2 on a few dozen rows: exact, or a near-tie of the reference
        assert a["ref_min_gap"] < (1e-4 if dtype == "fp32" else 2e-3), f"{name}/{dtype}: mismatch is not a near-tie: {a}"
    assert len(audits) <= 1


@pytest.mark.parametrize("dtype", ["fp32", "bf16x2", "bf16"])
def test_c3_beam5_full_fixture(dtype):
    """configs[2] beam search, width 5, 32 images (160 hypotheses in flight) against HF GenerationMixin on the same weights; every
    mode is compared with HF (bf16 beam was only compared with itself) and its match rate reported."""
    g = gu.load("c3_medium_tfm_beam5_full32")
    g = dict(g, emb_total=np.array(32))
    model, _, x = gpu_util.product_model(g, dtype)
    model.num_beams = 5
    ids = model.generate(image_embeddings=x.to(DEV), max_length=30, temperature=0.0).cpu().numpy()
    ref = g["ids"].astype(np.int64)
    L = min(ids.shape[1], ref.shape[1])
    same = [(ids.shape[1] == ref.shape[1]) and np.array_equal(ids[b], ref[b]) for b in range(32)]
    _report(test="c3_beam5_full", dtype=dtype, rows=32, rows_identical=int(sum(same)), shape=list(ids.shape), ref_shape=list(ref.shape),
            token_match=float((ids[:, :L] == ref[:, :L]).mean()))
    if dtype == "bf16":
        assert sum(same) >= 12, f"bf16 beam search: only {sum(same)}/32 hypotheses identical to HF fp32"
    else:
        assert sum(same) >= 31, f"{dtype} beam search: {sum(same)}/32 identical to HF"


@pytest.fixture(scope="module")
def c5_store():
    from gpt2_image_captioning_b200 import GpuFlatStore
    g = gu.load("c5_retrieval_full")
    n_img, n_cap = int(g["n_img"]), int(g["n_cap"])
    gen = torch.Generator().manual_seed(int(g["db_seed"]))
    img = torch.randn(n_img, 512, generator=gen)
    img /= img.norm(dim=-1, keepdim=True)
    cap = torch.randn(n_cap, 512, generator=gen)
    cap /= cap.norm(dim=-1, keepdim=True)
    names = [f"{i:012d}.jpg" for i in range(n_img)]
    store = GpuFlatStore(img, cap, names, [{"filename": names[min(j // 5, n_img - 1)]} for j in range(n_cap)], device=DEV)
    q = oc.synthetic_embeddings(1024, 512, 1)
    q[:4] = img[torch.from_numpy(g["exact_rows"].astype(np.int64))]
    return g, store, q


def _check_topk(idx, scores, ref_idx, ref_scores, ref_gap, what):
    np.testing.assert_allclose(scores, ref_scores, rtol=0, atol=2e-6, err_msg=what)
    bad = np.nonzero((idx != ref_idx).any(axis=1))[0]
    for b in bad:  # an fp32 order may differ from the float64 one only where two scores are closer than fp32 resolves
        assert ref_gap[b] < 1e-6, f"{what}: query {b} differs at a float64 gap of {ref_gap[b]:.3e}: {idx[b]} vs {ref_idx[b]}"
    return len(bad)


@pytest.mark.parametrize("exact", [False, True])
def test_c5_retrieval_at_full_size(c5_store, monkeypatch, exact):
    """configs[4] at its stated size: top-15 over the 118 287 x 512 image matrix and top-5 over the 591 753 x 512 caption matrix for
    1024 queries, against float64 inner products ordered (score desc, index asc) -- the tensor-core scan and the fp32 CUDA-core scan --
    then the reference's hit filter / caption-row selection / mean-add on the result (faiss_store.py:153-183,208-251)."""
    g, store, q = c5_store
    if exact:
        for ix in (store.image_index, store.caption_index):
            monkeypatch.setattr(ix, "hi", None)
    s, i = store.image_index.search_device(q.to(DEV), 15)
    n_img_ties = _check_topk(i.cpu().numpy(), s.cpu().numpy(), g["img_idx"].astype(np.int64), g["img_scores"], g["img_min_gap"], "image top-15")
    s, i = store.caption_index.search_device(q.to(DEV), 5)
    n_cap_ties = _check_topk(i.cpu().numpy(), s.cpu().numpy(), g["cap_idx"].astype(np.int64), g["cap_scores"], g["cap_min_gap"], "caption top-5")
    rows = store.retrieve_rows(q.to(DEV), top_i=5, top_k=5).cpu().numpy()
    ref_rows = g["rat_rows"].astype(np.int64)
    rows_bad = np.nonzero((rows != ref_rows).any(axis=1))[0]
    for b in rows_bad:
        assert g["img_min_gap"][b] < 1e-6
    assert (ref_rows[:4, 0] // 5 != g["exact_rows"]).all()  # the self-match of an exact database row is filtered (score > 0.9999)
    aug = store.retrieve_and_aggregate(q[:64].to(DEV), top_i=5, top_k=5).cpu().numpy()
    np.testing.assert_allclose(aug, g["aug64"], rtol=0, atol=2e-6)
    _report(test="c5_full", scan="fp32" if exact else "tensor-core", queries=1024, image_rows=118287, caption_rows=591753,
            image_near_tie_rows=n_img_ties, caption_near_tie_rows=n_cap_ties, rat_rows_mismatched=int(len(rows_bad)))


@pytest.mark.parametrize("dtype", ["bf16x2", "bf16"])
def test_c5_rat_captions_in_the_timed_dtypes(c5_store, dtype):
    """RAT end to end on the full-size database in the dtypes tools/bench_configs.py times: tokens of generate(db_store, ...) equal the
    tokens of the plain model on the augmented embeddings (retrieval is exact in every mode), and in bf16x2 they equal the fp32 engine's."""
    from gpt2_image_captioning_b200 import MLPMappingNetwork, RetrievalAugmentedTransformer
    from oracle.ref_harness import StubTokenizer
    g, store, q = c5_store
    gpt, mapper_ref = oc.build_modules(oc.ModelSpec())
    mapper = MLPMappingNetwork(prefix_length=10, embed_dim=512, gpt_dim=768)
    mapper.load_state_dict(mapper_ref.state_dict())
    rat = RetrievalAugmentedTransformer(512, 4, "mean", mapper, tokenizer=StubTokenizer(), gpt=gpt, engine_dtype=dtype).to(DEV)
    x = q[:256].to(DEV)
    got = rat.generate(store, 5, 5, x, max_length=30, temperature=0.0)
    aug = store.retrieve_and_aggregate(x, top_i=5, top_k=5)
    plain = rat._generate_on_engine(aug, 30, 0.0, 0.9)
    assert torch.equal(got, plain)
    rat.engine_dtype = "fp32"
    ref = rat._generate_on_engine(aug, 30, 0.0, 0.9)
    match = float((got == ref).all(dim=1).float().mean())
    _report(test="c5_rat_tokens", dtype=dtype, rows=int(x.shape[0]), caption_match_vs_fp32_engine=match)
    assert match >= (0.98 if dtype == "bf16x2" else 0.6), match


# ---- the reference's own objects through the boundary (VERDICT r1 missing item 4) -------------------------------------
class _RefShapedStore:
    """What `create_faiss_store` returns, as far as the model can tell: image_index / caption_index with search / reconstruct /
    ntotal, the two metadata lists and filename_to_caption_indices (src/database/faiss_store.py:16-52) -- numpy-backed, like the
    stand-in tests/golden/make_golden.py drives the unmodified reference with."""

    class _Flat:
        def __init__(self, m):
            self.m = np.ascontiguousarray(m, np.float32)
            self.ntotal, self.d = self.m.shape

        def search(self, q, k):
            return oc.flat_ip_search(self.m, q, k)

        def reconstruct(self, i):
            return self.m[int(i)]

    def __init__(self, img, cap, names, cap_meta):
        self.image_index, self.caption_index = self._Flat(img), self._Flat(cap)
        self.image_metadata, self.caption_metadata = names, cap_meta
        self.filename_to_caption_indices = {}
        for r, m in enumerate(cap_meta):
            self.filename_to_caption_indices.setdefault(m["filename"], []).append(r)

    def close(self):
        pass


def _small_db(D=64, n_img=1500, seed=33):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n_img, D)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    counts = rng.integers(0, 6, n_img)
    owner = np.repeat(np.arange(n_img), counts)
    cap = rng.standard_normal((len(owner), D)).astype(np.float32)
    names = [f"img_{i:06d}.jpg" for i in range(n_img)]
    starts = np.concatenate([[0], np.cumsum(counts)])
    return img, cap, names, [{"filename": names[o], "caption_id": j} for j, o in enumerate(owner)], starts


@pytest.mark.parametrize("aggregation", ["mean", "attention"])
def test_rat_generate_accepts_a_reference_faiss_store(aggregation):
    """`RetrievalAugmentedTransformer.generate(db_store=<reference FAISSStore>)` as src/eval.py:281-289 calls it: the store is
    uploaded once (cached per object), tokens equal the oracle's on the oracle's augmented embeddings; `forward` trains the
    aggregator (attention_proj must receive a gradient)."""
    from gpt2_image_captioning_b200 import MLPMappingNetwork, RetrievalAugmentedTransformer
    from oracle.ref_harness import StubTokenizer
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, mapper_ref = oc.build_modules(spec)
    oracle = oc.CaptionOracle(spec, gpt, mapper_ref)
    img, cap, names, meta, starts = _small_db()
    store = _RefShapedStore(img, cap, names, meta)
    mapper = MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
    mapper.load_state_dict(mapper_ref.state_dict())
    model = RetrievalAugmentedTransformer(64, 4, aggregation, mapper, tokenizer=StubTokenizer(), gpt=gpt, engine_dtype="fp32").to(DEV)
    x = oc.synthetic_embeddings(40, 64, 6)
    got = model.generate(db_store=store, top_k=10, top_i=4, image_embeddings=x.to(DEV), max_length=8, temperature=0.0)
    first = model._device_store(store)
    assert model._device_store(store) is first  # uploaded once
    if aggregation == "mean":
        want_aug, _ = oc.retrieve_and_aggregate(img, cap, lambda im: list(range(starts[im], starts[im + 1])), x.numpy(), top_i=4, top_k=10)
        assert torch.equal(got.cpu(), oracle.generate(torch.from_numpy(want_aug), 8, kv_cache=True))
    # training forward: differentiable pooling, constant retrieval
    model.train()
    cap_ids = torch.randint(0, 1000, (6, 5), device=DEV)
    xg = x[:6].to(DEV).requires_grad_(True)
    out = model(store, 4, 10, cap_ids, xg, attention_mask=torch.ones_like(cap_ids), labels=cap_ids)
    out.loss.backward()
    assert xg.grad is not None and float(xg.grad.abs().sum()) > 0
    if aggregation == "attention":
        gw = model.aggregator.attention_proj.weight.grad
        assert gw is not None and float(gw.abs().sum()) > 0, "attention_proj received no gradient"
    # the fused generate path and the module path compute the same augmentation
    with torch.no_grad():
        model.eval()
        fused = model._augment(store, x.to(DEV), 4, 10)
        modular = model.aggregator(x.to(DEV), model._retrieve_batch(store, x.to(DEV), 4, 10))
    assert torch.allclose(fused, modular, atol=2e-6)
    with pytest.raises(NotImplementedError):
        model.generate(db_store=object(), top_k=10, top_i=4, image_embeddings=x.to(DEV), max_length=4, temperature=0.0)


class _PlainReferenceModel(torch.nn.Module):
    """A module that only LOOKS like the reference's ImageCaptioningModel (.mapping_network / .gpt / .tokenizer / .task_prefix_embeds
    and its own slow generate) -- built from the oracle's modules, no product class involved."""

    def __init__(self, gpt, mapper, tokenizer):
        super().__init__()
        self.gpt, self.mapping_network, self.tokenizer = gpt, mapper, tokenizer
        self.task_prefix_embeds = None

    def generate(self, image_embeddings, max_length=50, temperature=1.0, top_p=0.9):
        raise AssertionError("accelerate() must have replaced this")


@pytest.mark.parametrize("mapper_kind", ["mlp", "transformer"])
def test_accelerate_attaches_the_engine_to_a_plain_reference_module(mapper_kind):
    """INTEGRATION.md option B: `accelerate(ref_model)` on an object that is NOT one of the product classes."""
    from gpt2_image_captioning_b200 import accelerate
    from oracle.ref_harness import StubTokenizer
    spec = (oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4) if mapper_kind == "mlp" else
            oc.ModelSpec(gpt="tiny", mapper="transformer", embed_dim=64, prefix_length=5, hidden_length=3, mapper_layers=2))
    gpt, mapper = oc.build_modules(spec)
    oracle = oc.CaptionOracle(spec, gpt, mapper)
    x = oc.synthetic_embeddings(12, 64, 8)
    want = oracle.generate(x, 9, kv_cache=True)
    import copy
    ref_model = _PlainReferenceModel(copy.deepcopy(gpt), copy.deepcopy(mapper), StubTokenizer()).to(DEV)
    fast = accelerate(ref_model, engine_dtype="fp32")
    assert fast is ref_model
    got = ref_model.generate(image_embeddings=x.to(DEV), max_length=9, temperature=0.0)
    assert got.device.type == "cuda" and torch.equal(got.cpu(), want)
    # weights edited through .data are invisible to the (data_ptr, _version) key: invalidate_engine() is the documented hook
    with torch.no_grad():
        ref_model.gpt.transformer.wte.weight.data.mul_(-1.0)
    ref_model.invalidate_engine()
    assert not torch.equal(ref_model.generate(image_embeddings=x.to(DEV), max_length=9, temperature=0.0).cpu(), want)


def test_eval_loop_on_the_gpu_dedupes_before_generating():
    """`generate_predictions` (the reference's eval loop, src/eval.py:199-224 / 232-308, with the image ids deduped BEFORE generation)
    through the real engine: same predictions as captioning every item and keeping the first per image, a fifth of the generate work;
    also the RAT variant with a reference-shaped store."""
    from gpt2_image_captioning_b200 import (ImageCaptioningModel, MLPMappingNetwork, RetrievalAugmentedTransformer, generate_predictions)
    from oracle.ref_harness import StubTokenizer

    class Tok(StubTokenizer):
        def batch_decode(self, ids, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) == self.eos_token_id)) for row in ids]

    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, mapper_ref = oc.build_modules(spec)
    mapper = MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
    mapper.load_state_dict(mapper_ref.state_dict())
    n_img = 37
    emb = oc.synthetic_embeddings(n_img, 64, 12)
    rng = np.random.default_rng(0)
    item_img = np.concatenate([np.repeat(np.arange(n_img), 5)])  # 5 caption items per image, as CocoDataset yields them
    rng.shuffle(item_img)
    batches = [{"image_id": torch.from_numpy(item_img[s:s + 16] + 1000), "image_embedding": emb[item_img[s:s + 16]]} for s in range(0, len(item_img), 16)]

    def reference_loop(model, **kw):  # src/eval.py:199-224 verbatim in structure: caption every item, keep the first per image id
        preds, seen = [], set()
        for b in batches:
            toks = model.generate(image_embeddings=b["image_embedding"].to(DEV), max_length=10, temperature=0.0, top_p=0.9, **kw)
            for i, c in zip(b["image_id"].tolist(), model.tokenizer.batch_decode(toks.cpu())):
                if i not in seen:
                    seen.add(i)
                    preds.append({"image_id": i, "caption": c})
        return preds

    model = ImageCaptioningModel(mapper, tokenizer=Tok(), gpt=gpt, engine_dtype="fp32").to(DEV)
    from gpt2_image_captioning_b200 import CaptionEngine
    n0 = CaptionEngine.launch_count()
    fast = generate_predictions(model, batches, batch_size=16, max_length=10, temperature=0.0, device=DEV)
    n1 = CaptionEngine.launch_count()
    slow = reference_loop(model)
    n2 = CaptionEngine.launch_count()
    assert fast == slow and len(fast) == n_img
    assert (n1 - n0) * 2 < (n2 - n1)  # ~5x fewer rows generated
    img, cap, names, meta, _ = _small_db()
    store = _RefShapedStore(img, cap, names, meta)
    rat = RetrievalAugmentedTransformer(64, 4, "mean", mapper, tokenizer=Tok(), gpt=gpt, engine_dtype="fp32").to(DEV)
    fast = generate_predictions(rat, batches, batch_size=16, max_length=10, temperature=0.0, device=DEV, db_store=store, top_k=10, top_i=4)
    slow = reference_loop(rat, db_store=store, top_k=10, top_i=4)
    assert fast == slow and len(fast) == n_img


def test_max_length_up_to_the_position_table():
    """The reference never feeds the last generated token back, so prefix + max_length - 1 == n_positions is legal."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork
    from oracle.ref_harness import StubTokenizer
    torch.manual_seed(0)
    gpt = GPT2LMHeadModel(GPT2Config(n_embd=128, n_layer=2, n_head=2, n_positions=16)).eval()
    mapper = MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
    model = ImageCaptioningModel(mapper, tokenizer=StubTokenizer(), gpt=gpt, engine_dtype="fp32").to(DEV)
    x = oc.synthetic_embeddings(3, 64, 2).to(DEV)
    out = model.generate(image_embeddings=x, max_length=13, temperature=0.0)  # 4 + 13 - 1 = 16 positions
    assert out.shape[0] == 3 and out.shape[1] <= 13
    with pytest.raises(Exception):
        model.generate(image_embeddings=x, max_length=14, temperature=0.0)


_SHARD_WORKER = r'''
import faulthandler, os, sys
faulthandler.enable()
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str({port}), RANK=str(rank), WORLD_SIZE=str(world))
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)  # the only communication of the path is the host gather of token ids
from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork, generate_for_embeddings
from oracle import captioner as oc
from oracle.ref_harness import StubTokenizer
spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
gpt, mapper_ref = oc.build_modules(spec)
mapper = MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
mapper.load_state_dict(mapper_ref.state_dict())
model = ImageCaptioningModel(mapper, tokenizer=StubTokenizer(), gpt=gpt, engine_dtype={dtype!r}).to(f"cuda:{{rank}}")
x = oc.synthetic_embeddings(1003, 64, 5)
ids = generate_for_embeddings(model, x, batch_size=128, max_length=12, device=f"cuda:{{rank}}", in_flight=2)
if rank == 0:
    assert ids.shape == (1003, 12) and ids.device.type == "cpu"
    dist.destroy_process_group()
    alone = generate_for_embeddings(model, x, batch_size=128, max_length=12, device="cuda:0", in_flight=1)
    assert torch.equal(ids, alone), "sharded + gathered ids differ from the 1-GPU ids"
    if {dtype!r} == "fp32":
        want = oc.CaptionOracle(spec, gpt.cpu(), mapper_ref).generate(x[500:540], 12, kv_cache=True)
        assert torch.equal(ids[500:540, : want.shape[1]], want)
    print("SHARDED_GPU_OK")
else:
    assert ids is None
    dist.destroy_process_group()
    print("SHARD_RANK_DONE", rank)
model.invalidate_engine()
torch.cuda.synchronize()
'''


@pytest.mark.parametrize("dtype", ["fp32", "bf16x2"])
def test_sharded_job_on_several_gpus(tmp_path, dtype):
    """north_star: "images shard independently across the GPUs of one box (no NCCL on the hot path; only a final host gather of
    captions)".  One process per GPU (up to 4), contiguous shards of 1003 rows, ragged batches, two batches in flight per GPU, ids
    gathered over gloo on rank 0: equal to the 1-GPU result (and to the oracle in fp32).  Skips on a single-GPU box."""
    import socket
    import subprocess
    import sys
    n = min(4, torch.cuda.device_count())
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER.format(root=root, port=port, dtype=dtype))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(n)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(n)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), ([p.returncode for p in procs], outs)
    assert "SHARDED_GPU_OK" in outs[0]
