"""Caption generation for an evaluation set with the duplicates removed BEFORE generation.

The reference's `generate_and_evaluate` / `generate_and_evaluate_rat` (src/eval.py:160-224, 227-300) iterate a `CocoDataset`
that yields one item per CAPTION (src/dataset.py:144-151: ~5 items per image), caption every item and only then keep the first
caption per `image_id` (:219-224) -- about 5x redundant generation on COCO (25 014 items for 5 000 val2017 images).  Greedy
decoding is deterministic and rows never interact, so generating once per first-seen image gives the same
`predictions` list.  `generate_predictions` is that loop: same arguments and the same `[{"image_id", "caption"}]` result in
first-seen order; metric computation (pycocoevalcap) stays with the caller, as a callback.

The embeddings file written by the reference's extractors (src/embeddings/clip.py:147-149, `{"filenames", "embeddings"}`) can
be captioned directly with `load_embeddings_pt` + `generate_for_embeddings`.
"""
from __future__ import annotations

from typing import Any, Callable, Iterable

import torch


def first_seen(image_ids: Iterable[int], seen: set[int]) -> list[int]:
    """Positions (within this batch) of the items whose image id has not been seen yet; `seen` is updated."""
    keep = []
    for pos, img_id in enumerate(image_ids):
        img_id = int(img_id)
        if img_id not in seen:
            seen.add(img_id)
            keep.append(pos)
    return keep


def generate_predictions(model, dataset, batch_size: int = 32, num_workers: int = 0, max_length: int = 50, temperature: float = 0.0,
                         top_p: float = 0.9, device: torch.device | str | None = None, db_store=None, top_k: int | None = None,
                         top_i: int | None = None, tokenizer=None, in_flight: int = 2) -> list[dict[str, Any]]:
    """One caption per distinct image of `dataset` (items with `image_id` and `image_embedding`), in first-seen order.

    `dataset` may be a torch Dataset (wrapped in a DataLoader exactly like src/eval.py:191-193) or any iterable of batches
    `{"image_id": LongTensor[b], "image_embedding": FloatTensor[b, E]}`.  With `db_store` the RAT signature is used
    (src/eval.py:281-289).  Unique rows are accumulated across loader batches so that `generate` always sees full batches of
    `batch_size` distinct images; `in_flight` of them run concurrently on the GPU (inflight.py), decoding to text stays on the
    calling thread."""
    from torch.utils.data import DataLoader, Dataset

    if temperature != 0:
        raise ValueError("dedupe-before-generate relies on deterministic (temperature = 0) decoding; with sampling every item's "
                         "caption is an independent draw and the reference keeps the FIRST item's")
    device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model = model.to(device)
    model.eval()
    batches = DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=num_workers) if isinstance(dataset, Dataset) else dataset
    tokenizer = tokenizer or getattr(dataset, "tokenizer", None) or model.tokenizer
    seen: set[int] = set()
    pend_ids: list[int] = []
    pend_emb: list[torch.Tensor] = []
    work: list[tuple[list[int], torch.Tensor]] = []  # full batches of distinct images, in first-seen order

    def cut(n: int) -> None:
        emb = torch.cat(pend_emb, dim=0)
        work.append((pend_ids[:n], emb[:n]))
        del pend_ids[:n]
        pend_emb.clear()
        if emb.shape[0] > n:
            pend_emb.append(emb[n:])

    for batch in batches:  # embeddings are precomputed: one pass over the loader costs I/O only (10 MB for val2017)
        ids = batch["image_id"].tolist() if torch.is_tensor(batch["image_id"]) else list(batch["image_id"])
        keep = first_seen(ids, seen)
        if keep:
            pend_ids.extend(int(ids[p]) for p in keep)
            pend_emb.append(batch["image_embedding"][keep])
        while len(pend_ids) >= batch_size:
            cut(batch_size)
    if pend_ids:
        cut(len(pend_ids))

    def run(item) -> torch.Tensor:
        kw = dict(image_embeddings=item[1].to(device), max_length=max_length, temperature=temperature, top_p=top_p)
        if db_store is not None:
            kw.update(db_store=db_store, top_k=top_k, top_i=top_i)
        with torch.no_grad():
            return model.generate(**kw).to("cpu")

    from .inflight import map_batches
    predictions: list[dict[str, Any]] = []
    for (ids, _), tokens in zip(work, map_batches(run, work, in_flight)):  # `in_flight` batches run concurrently on the GPU
        captions = tokenizer.batch_decode(tokens, skip_special_tokens=True)
        predictions.extend({"image_id": i, "caption": c} for i, c in zip(ids, captions))
    return predictions


def generate_and_evaluate(model, dataset, annotations_path: str, evaluate_fn: Callable[[list[dict[str, Any]], str], Any], **kwargs):
    """Drop-in shape of src/eval.py:160-224: (predictions, metrics), with the metric code supplied by the caller
    (`evaluate_fn = src.eval.evaluate_captions` in the reference repository)."""
    predictions = generate_predictions(model, dataset, **kwargs)
    return predictions, evaluate_fn(predictions, annotations_path)


def load_embeddings_pt(path: str) -> tuple[list[str], torch.Tensor]:
    """The extractors' output file (src/embeddings/clip.py:147-149): {"filenames": [...], "embeddings": FloatTensor[N, E]}."""
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(blob, dict) or "embeddings" not in blob or "filenames" not in blob:
        raise ValueError(f"{path}: expected a dict with 'filenames' and 'embeddings' (src/embeddings/clip.py:147-149)")
    emb = torch.as_tensor(blob["embeddings"], dtype=torch.float32)
    names = list(blob["filenames"])
    if emb.dim() != 2 or emb.shape[0] != len(names):
        raise ValueError(f"{path}: {len(names)} filenames for an embeddings tensor of shape {tuple(emb.shape)}")
    return names, emb


def generate_for_embeddings(model, embeddings: torch.Tensor, batch_size: int = 1024, max_length: int = 50, device=None,
                            in_flight: int = 2, group=None, **generate_kw) -> torch.Tensor:
    """int64 [N, max_length] token ids (EOS-padded) for a whole embeddings matrix, sharded over the ranks of an initialised
    process group if there is one (sharding.generate_sharded; other ranks get None).  `in_flight` batches run concurrently per GPU
    (inflight.py: +10 % captions/s at 2, each extra slot is one more engine context -- workspace, KV cache, CUDA graph -- on the same
    packed weights); ids do not depend on it.  `group`: the process group of the final host gather (e.g. a gloo group next to an NCCL
    default group, so that the ids travel host to host)."""
    from .sharding import generate_sharded

    device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model = model.to(device).eval()
    fn = lambda x: model.generate(image_embeddings=x.to(device), max_length=max_length, temperature=0.0, **generate_kw)
    return generate_sharded(fn, embeddings, max_length, batch_size, int(model.tokenizer.eos_token_id), group=group, in_flight=in_flight)
