"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see oracle/ref_harness.py).
    python tests/golden/make_golden.py [case ...]      # default: all fast cases
Cases:  c1  tiny_eos  tiny_tfm  tiny_task  c2_1024  c3  c4  rat
Each .npz stores the inputs' pins (spec, seeds, weight fingerprint), the reference's token ids
and, for a few rows, its fp32 last-position logits at step 0.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import captioner as oc  # noqa: E402
from oracle import ref_harness  # noqa: E402


def build_reference_model(ref, spec: oc.ModelSpec, eos: int = oc.EOS_TOKEN_ID, task_prefix: torch.Tensor | None = None,
                          eos_row_scale: float = 1.0):
    """Same RNG order as oracle.captioner.build_modules, but with the reference's own classes."""
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(spec.seed)
    gpt = GPT2LMHeadModel(GPT2Config(**spec.dims))
    d = spec.dims["n_embd"]
    if spec.mapper == "mlp":
        mapper = ref.MLPMappingNetwork(prefix_length=spec.prefix_length, embed_dim=spec.embed_dim, gpt_dim=d)
    else:
        mapper = ref.TransformerMappingNetwork(embed_dim=spec.embed_dim, gpt_dim=d, prefix_length=spec.prefix_length,
                                               hidden_length=spec.hidden_length, num_layers=spec.mapper_layers)
    if eos_row_scale != 1.0:  # see case_tiny_eos
        with torch.no_grad():
            gpt.transformer.wte.weight[eos] *= eos_row_scale
    tok = ref_harness.StubTokenizer()
    tok.eos_token_id = eos
    model = ref.ImageCaptioningModel(mapping_network=mapper, tokenizer=tok, gpt=gpt)
    if task_prefix is not None:  # what prefix_task_prompt would have produced (src/models.py:219-235)
        model.task_prefix_embeds = torch.nn.Parameter(task_prefix.clone())
    return model.eval()


def step0_logits(model, x):
    with torch.no_grad():
        p = model.mapping_network(x)
        if model.task_prefix_embeds is not None:
            p = torch.cat((p, model.task_prefix_embeds.unsqueeze(0).expand(x.shape[0], -1, -1)), dim=1)
        return model.gpt(inputs_embeds=p).logits[:, -1, :].float().numpy()


def spec_dict(spec: oc.ModelSpec) -> dict:
    return {f"spec_{k}": np.array(v) for k, v in spec.__dict__.items()}


def save(name: str, spec: oc.ModelSpec, model, **arrays):
    fp = oc.weight_fingerprint(model.gpt, model.mapping_network)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **spec_dict(spec), **{f"fp_{k}": np.array(v) for k, v in fp.items()}, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)", flush=True)


def greedy_case(ref, name, spec, n_rows, max_length, batch, emb_seed=1, emb_total=None, logit_rows=2, eos=oc.EOS_TOKEN_ID,
                task_prefix=None, eos_row_scale=1.0):
    model = build_reference_model(ref, spec, eos, task_prefix, eos_row_scale)
    x = oc.synthetic_embeddings(emb_total or n_rows, spec.embed_dim, emb_seed)[:n_rows]
    t0 = time.time()
    ids, lens = [], []
    for i in range(0, n_rows, batch):
        out = model.generate(image_embeddings=x[i:i + batch], max_length=max_length, temperature=0.0, top_p=0.9)
        lens.append(out.shape[1])
        pad = torch.full((out.shape[0], max_length - out.shape[1]), -1, dtype=out.dtype)
        ids.append(torch.cat((out, pad), dim=1))
        print(f"  {name}: rows {i + out.shape[0]}/{n_rows}  L_gen={out.shape[1]}  {time.time() - t0:.1f}s", flush=True)
    ids = torch.cat(ids).numpy().astype(np.int32)
    extra = {}
    if task_prefix is not None:
        extra["task_prefix"] = task_prefix.numpy()
    save(name, spec, model, ids=ids, batch_lens=np.array(lens), batch=np.array(batch), n_rows=np.array(n_rows),
         max_length=np.array(max_length), emb_seed=np.array(emb_seed), emb_total=np.array(emb_total or n_rows),
         eos=np.array(eos), eos_row_scale=np.array(eos_row_scale), logits0=step0_logits(model, x[:logit_rows]), ref_seconds=np.array(time.time() - t0), **extra)
    return model, x, ids


def case_c1(ref):
    """BASELINE.json config 1: GPT-2 small + MLP mapper, 512-d, P=10, greedy 30, batch 64, CPU fp32."""
    greedy_case(ref, "c1_small_mlp_b64", oc.ModelSpec(), 64, 30, 64)


def case_c2_1024(ref):
    """First 1024 rows of config 2's 5 000-row embedding set (same model as c1)."""
    greedy_case(ref, "c2_small_mlp_first1024", oc.ModelSpec(), 1024, 30, 128, emb_total=5000)


def case_tiny_eos(ref):
    """Random-init GPT-2 never emits EOS, so to pin the EOS / early-break / trim rules (src/models.py:390-391,
    453-460) the tied wte row of the EOS token is scaled x6 after the seeded construction: its logit then wins at
    different steps in different rows.  Batches of 4 => some batches stop early, others run to max_length."""
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    greedy_case(ref, "tiny_mlp_eos", spec, 48, 16, 4, logit_rows=4, eos_row_scale=6.0)


def case_tiny_tfm(ref):
    spec = oc.ModelSpec(gpt="tiny", mapper="transformer", embed_dim=64, prefix_length=5, hidden_length=3, mapper_layers=2)
    greedy_case(ref, "tiny_tfm", spec, 16, 12, 16, logit_rows=4)


def case_tiny_task(ref):
    """Task-prompt prefix appended after the image prefix (src/models.py:364-375)."""
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    g = torch.Generator().manual_seed(7)
    task = torch.randn(3, 128, generator=g) * 0.02
    greedy_case(ref, "tiny_mlp_task", spec, 16, 10, 16, logit_rows=4, task_prefix=task)


def case_c3(ref):
    """Config 3: GPT-2 medium + 8-layer transformer mapper, P=40 (greedy through the reference; beam-5 through HF
    GenerationMixin on model.gpt, SURVEY.md 8(a) A9, since the reference has no beam search)."""
    spec = oc.ModelSpec(gpt="medium", mapper="transformer", embed_dim=512, prefix_length=40, hidden_length=10, mapper_layers=8)
    model, x, _ = greedy_case(ref, "c3_medium_tfm", spec, 8, 30, 8)
    with torch.no_grad():
        p = model.mapping_network(x[:4])
        beam = model.gpt.generate(inputs_embeds=p, num_beams=5, do_sample=False, max_new_tokens=30, early_stopping=False,
                                  length_penalty=1.0, num_return_sequences=1, eos_token_id=oc.EOS_TOKEN_ID,
                                  pad_token_id=oc.EOS_TOKEN_ID)
    path = os.path.join(HERE, "c3_medium_tfm_beam5.npz")
    np.savez_compressed(path, **spec_dict(spec), ids=beam.numpy().astype(np.int32), n_rows=np.array(4), num_beams=np.array(5),
                        max_length=np.array(30))
    print("wrote", path, beam.shape)


def case_tiny_beam(ref):
    spec = oc.ModelSpec(gpt="tiny", embed_dim=64, prefix_length=4)
    model = build_reference_model(ref, spec)
    x = oc.synthetic_embeddings(8, 64, 1)
    with torch.no_grad():
        p = model.mapping_network(x)
        beam = model.gpt.generate(inputs_embeds=p, num_beams=5, do_sample=False, max_new_tokens=12, early_stopping=False,
                                  length_penalty=1.0, num_return_sequences=1, eos_token_id=oc.EOS_TOKEN_ID,
                                  pad_token_id=oc.EOS_TOKEN_ID)
    path = os.path.join(HERE, "tiny_mlp_beam5.npz")
    np.savez_compressed(path, **spec_dict(spec), ids=beam.numpy().astype(np.int32), n_rows=np.array(8), num_beams=np.array(5),
                        max_length=np.array(12))
    print("wrote", path, beam.shape)


def case_c4(ref):
    """Config 4: GPT-2 large + MLP mapper on 1024-d embeddings, greedy 30."""
    spec = oc.ModelSpec(gpt="large", embed_dim=1024, prefix_length=10)
    greedy_case(ref, "c4_large_mlp", spec, 4, 30, 4)


class NumpyFlatIP:
    """Stand-in for faiss.IndexFlatIP (not installable here): exact inner product, descending, lowest index first
    on ties -- the duck-typed surface the reference touches (search / reconstruct), SURVEY.md 8(b)."""

    def __init__(self, mat):
        self.mat = np.ascontiguousarray(mat, np.float32)

    def search(self, q, k):
        return oc.flat_ip_search(self.mat, q, k)

    def reconstruct(self, i):
        return self.mat[i]


def case_rat(ref, fstore):
    """Drive the reference's own retrieve_images_by_vector_similarity / get_caption_embeddings / RetrievalAggregator
    (src/database/faiss_store.py:132-251, src/models.py:589-625,655-695) through the stand-in index."""
    rng = np.random.default_rng(2)
    n_img, D = 2000, 512
    img = rng.standard_normal((n_img, D)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    # ragged captions per image: 0..6 (0 exercises the zero-padding path with few hits)
    counts = rng.integers(0, 7, n_img)
    owner = np.repeat(np.arange(n_img), counts)
    cap = rng.standard_normal((len(owner), D)).astype(np.float32)  # caption embeddings are NOT normalised (word_embeddings.py:139-140)
    names = [f"img_{i:06d}.jpg" for i in range(n_img)]
    store = fstore.FAISSStore(NumpyFlatIP(img), NumpyFlatIP(cap), names, [{"filename": names[o], "caption_id": j} for j, o in enumerate(owner)])
    q = oc.synthetic_embeddings(40, D, 1).numpy()
    q[:6] = img[[5, 17, 100, 101, 1999, 0]]  # exact DB rows -> the > 0.9999 self-match filter
    q[6] = 0.5 * (img[3] + img[4]) / np.linalg.norm(0.5 * (img[3] + img[4]))
    out = {}
    for (top_k, top_i) in [(10, 4), (20, 6), (5, 1)]:
        res = fstore.retrieve_images_by_vector_similarity(store, q, top_i)
        files = [[f for f, _ in r] for r in res]
        emb = fstore.get_caption_embeddings(store, top_k, files, embed_dim=D)
        agg = ref.RetrievalAggregator(D, "mean")(torch.from_numpy(q), torch.from_numpy(emb)).numpy()
        out[f"ret_k{top_k}_i{top_i}"] = emb.astype(np.float32)
        out[f"aug_k{top_k}_i{top_i}"] = agg.astype(np.float32)
        out[f"hits_k{top_k}_i{top_i}"] = np.array([[names.index(f) for f in fl] + [-1] * (top_i - len(fl)) for fl in files], np.int64)
    path = os.path.join(HERE, "rat_retrieval.npz")
    np.savez_compressed(path, db_seed=np.array(2), n_img=np.array(n_img), counts=counts, q=q, **out)
    print("wrote", path)


FAST = ["c1", "tiny_eos", "tiny_tfm", "tiny_task", "tiny_beam", "rat"]

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    ref, fstore = ref_harness.import_reference()
    cases = sys.argv[1:] or FAST
    for c in cases:
        print("==", c, flush=True)
        fn = globals()["case_" + c]
        fn(ref, fstore) if c == "rat" else fn(ref)
