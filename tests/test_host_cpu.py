"""CPU-only checks (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares, the host-side
mirror keeps the reference's API surface, the product path fails loudly without CUDA, and the N>1 sharding logic works
over gloo with world_size 2."""
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import captioner as oc
from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    lib = os.path.join(ROOT, "gpt2_image_captioning_b200", "libgic_b200.so")
    if not os.path.isfile(lib):
        import __graft_entry__
        __graft_entry__.build()
    return lib


def test_library_exports_every_declared_symbol(built_library):
    from gpt2_image_captioning_b200 import _capi
    header = open(os.path.join(ROOT, "include", "gic_b200.h")).read()
    declared = set(re.findall(r"GIC_API\s+[\w\s\*]+?\b(gic_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    L = _capi.lib()  # dlopen; raises if a symbol is missing
    for name in declared:
        assert hasattr(L, name)
    assert L.gic_abi_version() == _capi.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", built_library], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gic_\w+)", nm))
    assert declared <= exported


def test_library_is_sm100a_tensor_core_code(built_library):
    sass = subprocess.run(["cuobjdump", "-sass", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma / TMA / tcgen05.ld (B200_PROFILING.md)
        assert mnemonic in sass, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_path_fails_loudly_without_cuda():
    from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork, GpuFlatStore, _capi
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, mref = oc.build_modules(spec)
    mapper = MLPMappingNetwork(4, 64, 128)
    model = ImageCaptioningModel(mapper, tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.generate(image_embeddings=torch.zeros(2, 64), max_length=3, temperature=0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        GpuFlatStore(np.zeros((4, 8), np.float32), np.zeros((4, 8), np.float32), ["a"] * 4, [{"filename": "a"}] * 4)
    assert _capi.lib().gic_device_check() != 0  # no device -> error code, not a crash
    assert b"" != _capi.lib().gic_last_error()


def test_no_product_import_of_the_oracle():
    pkg = os.path.join(ROOT, "gpt2_image_captioning_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_mirror_classes_keep_reference_surface():
    import gpt2_image_captioning_b200 as pkg
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, mref = oc.build_modules(spec)
    mapper = pkg.MLPMappingNetwork(prefix_length=4, embed_dim=64, gpt_dim=128)
    assert set(mapper.state_dict()) == set(mref.state_dict())
    tfm = pkg.TransformerMappingNetwork(embed_dim=64, gpt_dim=128, prefix_length=5, hidden_length=3, num_layers=2)
    assert tfm(torch.zeros(2, 64)).shape == (2, 5, 128)
    model = pkg.RetrievalAugmentedTransformer(64, 4, "attention", mapper, tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    keys = set(model.state_dict())
    assert any(k.startswith("mapping_network.model.0.") for k in keys) and any(k.startswith("gpt.transformer.h.0.") for k in keys)
    assert {"aggregator.attention_proj.weight", "aggregator.attention_proj.bias"} <= keys
    sig = inspect.signature(pkg.ImageCaptioningModel.generate)
    assert list(sig.parameters)[:5] == ["self", "image_embeddings", "max_length", "temperature", "top_p"]
    assert (sig.parameters["max_length"].default, sig.parameters["temperature"].default, sig.parameters["top_p"].default) == (50, 1.0, 0.9)
    rsig = inspect.signature(pkg.RetrievalAugmentedTransformer.generate)
    assert list(rsig.parameters)[:5] == ["self", "db_store", "top_k", "top_i", "image_embeddings"]
    # training forward stays plain PyTorch and works on the CPU
    out = pkg.ImageCaptioningModel(mapper, tokenizer=ref_harness.StubTokenizer(), gpt=gpt).forward(
        caption_token_ids=torch.randint(0, 100, (2, 6)), image_embeddings=torch.randn(2, 64),
        attention_mask=torch.ones(2, 6, dtype=torch.long), labels=torch.randint(0, 100, (2, 6)))
    assert out.logits.shape == (2, 4 + 6, 50257) and out.loss is not None


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present")
def test_mirror_matches_live_reference_signatures_and_keys():
    import gpt2_image_captioning_b200 as pkg
    ref, _ = ref_harness.import_reference()
    for cls in ("MLPMappingNetwork", "TransformerMappingNetwork", "ImageCaptioningModel", "RetrievalAggregator",
                "RetrievalAugmentedTransformer"):
        r, m = getattr(ref, cls), getattr(pkg, cls)
        rp = [p for p in inspect.signature(r.__init__).parameters.values()]
        mp = [p for p in inspect.signature(m.__init__).parameters.values() if p.kind != p.KEYWORD_ONLY]
        assert [p.name for p in rp] == [p.name for p in mp], cls
        for meth in ("generate", "forward", "generate_captions", "save_parameters", "load_saved_parameters", "_retrieve_batch"):
            if hasattr(r, meth):
                assert list(inspect.signature(getattr(r, meth)).parameters) == list(inspect.signature(getattr(m, meth)).parameters), (cls, meth)
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, _ = oc.build_modules(spec)
    a = ref.ImageCaptioningModel(ref.MLPMappingNetwork(4, 64, 128), tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    b = pkg.ImageCaptioningModel(pkg.MLPMappingNetwork(4, 64, 128), tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())  # a reference checkpoint loads into the mirror
    x = torch.randn(3, 64)
    ids = torch.randint(0, 1000, (3, 5))
    assert torch.allclose(a.forward(ids, x).logits, b.forward(ids, x).logits)


def test_save_and_load_parameters_roundtrip(tmp_path):
    import gpt2_image_captioning_b200 as pkg
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    gpt, _ = oc.build_modules(spec)
    m1 = pkg.ImageCaptioningModel(pkg.MLPMappingNetwork(4, 64, 128), tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    path = str(tmp_path / "ckpt.pt")
    m1.save_parameters(path)
    saved = torch.load(path)
    assert saved and all(not k.startswith("gpt.") for k in saved)  # frozen GPT-2 is not saved (src/models.py:489-519)
    m2 = pkg.ImageCaptioningModel(pkg.MLPMappingNetwork(4, 64, 128), tokenizer=ref_harness.StubTokenizer(), gpt=gpt)
    m2.load_saved_parameters(path)
    assert torch.equal(m1.mapping_network.model[0].weight, m2.mapping_network.model[0].weight)
    torch.save({"bogus": torch.zeros(1)}, path)
    with pytest.raises(ValueError):
        m2.load_saved_parameters(path)


# ---- sharding ---------------------------------------------------------------------------------------------------------
def test_shard_range_covers_everything_in_order():
    from gpt2_image_captioning_b200 import shard_range
    for n in (0, 1, 7, 5000, 118287):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard_range(118287, 0, 8) == (0, 14786) and shard_range(118287, 7, 8)[1] == 118287
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_generate_batches_pads_with_eos():
    from gpt2_image_captioning_b200.sharding import generate_batches
    calls = []

    def fake(xb):  # a "model.generate" whose L_gen varies per call
        calls.append(xb.shape[0])
        L = 2 if len(calls) == 1 else 4
        return torch.arange(xb.shape[0] * L).view(xb.shape[0], L)

    out = generate_batches(fake, torch.zeros(5, 8), max_length=4, batch_size=3, eos_token_id=9)
    assert calls == [3, 2] and out.shape == (5, 4)
    assert out[0].tolist() == [0, 1, 9, 9] and out[4].tolist() == [4, 5, 6, 7]


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from gpt2_image_captioning_b200.sharding import generate_sharded
from oracle import captioner as oc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
o = oc.CaptionOracle(spec)
x = oc.synthetic_embeddings(11, 64, 1)
out = generate_sharded(lambda xb: o.generate(xb, 5, kv_cache=True), x, max_length=5, batch_size=4)
if dist.get_rank() == 0:
    want = o.generate(x, 5, kv_cache=True)
    assert out.shape == (11, 5) and torch.equal(out, want), (out, want)
    print("SHARDED_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_sharded_generate_world_size_2_gloo(tmp_path):
    """Two CPU processes over gloo, each captioning its contiguous shard (the oracle stands in for the per-rank GPU
    worker -- this tests the host-side split / gather / ordering only): the gathered result equals the single-process one."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARDED_OK" in outs[0]


class _FakeTok:
    eos_token_id = 50256

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) == self.eos_token_id)) for row in ids]


class _FakeCaptioner(torch.nn.Module):
    """deterministic, row-independent stand-in for ImageCaptioningModel.generate (counts the rows it is asked to caption)"""

    def __init__(self):
        super().__init__()
        self.tokenizer = _FakeTok()
        self.rows = 0
        self.batches = []

    def generate(self, image_embeddings, max_length=50, temperature=1.0, top_p=0.9, **kw):
        self.rows += image_embeddings.shape[0]
        self.batches.append(image_embeddings.shape[0])
        base = (image_embeddings[:, 0] * 1000).round().long()
        return torch.stack([base + t for t in range(max_length)], dim=1)


def test_dedupe_before_generate_matches_the_reference_loop(tmp_path):
    """generate_predictions == the reference's generate_and_evaluate loop (src/eval.py:199-224: caption every item, keep the first
    per image_id), with one generated row per distinct image instead of one per caption item."""
    from gpt2_image_captioning_b200 import generate_predictions, load_embeddings_pt
    g = torch.Generator().manual_seed(0)
    n_img = 23
    emb = torch.rand(n_img, 8, generator=g)
    owner = torch.tensor([i for i in range(n_img) for _ in range(1 + i % 5)])  # 1..5 caption items per image, grouped like CocoDataset
    items = [{"image_id": int(100 + o), "image_embedding": emb[o]} for o in owner]

    class DS(torch.utils.data.Dataset):
        tokenizer = _FakeTok()

        def __len__(self):
            return len(items)

        def __getitem__(self, i):
            return items[i]

    # the reference loop, restated
    ref_model, want, seen = _FakeCaptioner(), [], set()
    for s in range(0, len(items), 7):
        chunk = items[s:s + 7]
        caps = ref_model.tokenizer.batch_decode(ref_model.generate(torch.stack([c["image_embedding"] for c in chunk]), max_length=6, temperature=0.0))
        for c, cap in zip(chunk, caps):
            if c["image_id"] not in seen:
                seen.add(c["image_id"])
                want.append({"image_id": c["image_id"], "caption": cap})
    model = _FakeCaptioner()
    got = generate_predictions(model, DS(), batch_size=7, max_length=6, temperature=0.0, device="cpu")
    assert got == want
    assert model.rows == n_img and ref_model.rows == len(items)
    assert sorted(model.batches) == [2, 7, 7, 7]  # full batches of distinct images + the remainder (two in flight: any order)
    one = _FakeCaptioner()
    assert generate_predictions(one, DS(), batch_size=7, max_length=6, temperature=0.0, device="cpu", in_flight=1) == want
    assert one.batches == [7, 7, 7, 2]
    with pytest.raises(ValueError, match="deterministic"):
        generate_predictions(model, DS(), temperature=1.0, device="cpu")
    # the extractors' .pt format
    path = tmp_path / "emb.pt"
    torch.save({"filenames": [f"{i}.jpg" for i in range(n_img)], "embeddings": emb}, path)
    names, e2 = load_embeddings_pt(str(path))
    assert names[3] == "3.jpg" and torch.equal(e2, emb)
    torch.save({"embeddings": emb}, path)
    with pytest.raises(ValueError):
        load_embeddings_pt(str(path))


def test_map_batches_order_slots_and_errors():
    """inflight.map_batches: results in batch order, batch i on slot i % in_flight, in_flight = 1 stays on the calling thread,
    worker exceptions reach the caller."""
    import threading
    from gpt2_image_captioning_b200.inflight import current_slot, map_batches
    main = threading.get_ident()
    seen = map_batches(lambda b: (b, current_slot(), threading.get_ident()), list(range(7)), in_flight=3)
    assert [s[0] for s in seen] == list(range(7)) and [s[1] for s in seen] == [i % 3 for i in range(7)]
    assert all(s[2] != main for s in seen)
    seq = map_batches(lambda b: (b, current_slot(), threading.get_ident()), list(range(4)), in_flight=1)
    assert all(s[1] == 0 and s[2] == main for s in seq)
    assert map_batches(lambda b: b, [], in_flight=2) == [] and current_slot() == 0

    def boom(b):
        if b == 3:
            raise KeyError("batch 3")
        return b
    with pytest.raises(KeyError, match="batch 3"):
        map_batches(boom, list(range(6)), in_flight=2)


def test_generate_batches_in_flight_equals_sequential():
    from gpt2_image_captioning_b200.sharding import generate_batches
    x = torch.arange(23 * 4, dtype=torch.float32).reshape(23, 4)

    def fake_generate(e):  # [b, L_gen] with a batch-dependent L_gen, like a trimmed reference batch
        L = 3 + int(e[0, 0].item()) % 3
        return (e[:, :1].long() + torch.arange(L)).contiguous()
    a = generate_batches(fake_generate, x, 6, 5, in_flight=1)
    b = generate_batches(fake_generate, x, 6, 5, in_flight=3)
    assert a.shape == (23, 6) and torch.equal(a, b)


def test_rejection_rule_of_the_sampler_is_the_reference_nucleus():
    """sample_top_p_kernel (csrc/lmhead.cu) draws from the FULL softmax and accepts a token iff the probability mass strictly above
    its own probability is <= top_p * S.  That accepted set must be exactly the set the reference keeps (sort, cumulative softmax,
    keep everything up to and including the first token whose cumulative probability exceeds top_p, src/models.py:413-432) -- then
    rejection sampling IS sampling from the reference's renormalised nucleus.  Checked on random rows, peaked rows and rows with ties."""
    import numpy as np
    import torch

    rng = np.random.default_rng(0)
    rows = [rng.normal(size=500) * s for s in (0.5, 2.0, 6.0)]
    tied = rng.normal(size=500) * 2.0
    tied[100:140] = tied[100]  # a plateau of equal logits that straddles some thresholds
    rows.append(tied)
    for z_np in rows:
        for top_p in (0.05, 0.5, 0.9, 0.999):
            for temperature in (0.7, 1.0):
                z = torch.from_numpy(z_np / temperature).float().unsqueeze(0)
                sl, si = torch.sort(z, descending=True)
                cp = torch.cumsum(torch.softmax(sl, dim=-1), dim=-1)
                rem = cp > top_p
                rem[:, 1:] = rem[:, :-1].clone()
                rem[:, 0] = False
                kept_ref = ~rem.scatter(1, si, rem)[0].numpy()
                p = np.exp((z[0].double().numpy() - z[0].double().numpy().max()))
                S = p.sum()
                above = np.array([p[p > pi].sum() for pi in p])
                kept_rule = above <= top_p * S
                # ties: the reference's sort order decides which of several EQUAL probabilities fall behind the cut, the rule keeps or
                # drops a plateau as a whole; away from a plateau at the cut the two sets are identical
                differ = kept_ref != kept_rule
                if differ.any():
                    cut = p[differ]
                    assert np.allclose(cut, cut[0]) and (np.isclose(p, cut[0]).sum() > 1), (top_p, temperature, int(differ.sum()))
                    assert kept_rule[differ].all()  # the plateau is kept whole: a superset of the reference's set by equal-probability tokens only
                assert kept_rule[np.argmax(p)]


def test_host_logic_properties():
    """Property-based checks (hypothesis) of the host logic every job goes through: the rank shards tile the rows in order and are
    balanced; generate_batches reassembles any split into the rows' own results whatever the batch size or the number of batches in
    flight, right-padded with EOS; the dedupe of the eval loop keeps the first position of every image id."""
    from hypothesis import given, settings, strategies as st
    from gpt2_image_captioning_b200 import shard_range
    from gpt2_image_captioning_b200.evaluation import first_seen
    from gpt2_image_captioning_b200.sharding import generate_batches

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 200000), st.integers(1, 16))
    def shards(n, world):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 70), st.integers(1, 33), st.integers(1, 3), st.integers(1, 9))
    def batches(n, batch, in_flight, max_length):
        x = torch.arange(n, dtype=torch.float32).unsqueeze(1)

        def fake_generate(e):  # row r -> r's own tokens, trimmed to a length that depends on the batch (as L_gen does)
            L = 1 + int(e[:, 0].max().item()) % max_length
            return (e[:, :1].long() % 1000).expand(-1, L).contiguous()

        out = generate_batches(fake_generate, x, max_length, batch, eos_token_id=50256, in_flight=in_flight)
        assert out.shape == (n, max_length)
        for s in range(0, n, batch):
            L = 1 + (min(n, s + batch) - 1) % max_length
            rows = out[s:s + batch]
            assert bool((rows[:, :L] == (torch.arange(s, min(n, s + batch)) % 1000).unsqueeze(1)).all())
            assert bool((rows[:, L:] == 50256).all())

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.lists(st.integers(0, 30), max_size=12), max_size=8))
    def dedupe(id_batches):
        seen: set = set()
        kept = []
        for ids in id_batches:
            kept.extend(ids[p] for p in first_seen(ids, seen))
        flat = [i for ids in id_batches for i in ids]
        assert kept == list(dict.fromkeys(flat))

    shards()
    batches()
    dedupe()
