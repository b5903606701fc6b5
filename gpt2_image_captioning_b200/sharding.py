"""Image sharding across the GPUs of one box: rows of `image_embeddings` never interact, so each
rank captions a contiguous shard with its own weight replica and the token ids are gathered on the host in rank order.
No collective on the hot path -- the only communication is the final gather of int64 ids (28 MB for 118 287 x 30).
"""
from __future__ import annotations

from typing import Callable

import torch

EOS_PAD = 50256


def shard_range(n_rows: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced split; the remainder goes to the low ranks.  [start, stop) of `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def generate_batches(generate_fn: Callable[[torch.Tensor], torch.Tensor], embeddings: torch.Tensor, max_length: int,
                     batch_size: int, eos_token_id: int = EOS_PAD, in_flight: int = 1) -> torch.Tensor:
    """Run `generate_fn` (one reference-style `model.generate` call per batch) over `embeddings` and return int64
    [n, max_length] on the CPU, each batch's [b, L_gen] right-padded with EOS (what every finished row would have
    produced had the loop continued, src/models.py:458-460).  `in_flight` > 1 keeps that many batches running
    concurrently on the GPU (inflight.map_batches: one stream + engine slot + host thread each); same ids."""
    from .inflight import map_batches

    out = torch.full((embeddings.shape[0], max_length), eos_token_id, dtype=torch.int64)
    starts = list(range(0, embeddings.shape[0], batch_size))
    results = map_batches(lambda s: generate_fn(embeddings[s:s + batch_size]).to("cpu"), starts, in_flight)
    for s, ids in zip(starts, results):
        out[s:s + ids.shape[0], : ids.shape[1]] = ids
    return out


def generate_sharded(generate_fn: Callable[[torch.Tensor], torch.Tensor], embeddings: torch.Tensor, max_length: int,
                     batch_size: int, eos_token_id: int = EOS_PAD, group=None, dst: int = 0, in_flight: int = 1) -> torch.Tensor | None:
    """Every rank captions rows shard_range(n, rank, world) of the SAME `embeddings` tensor; rank `dst` returns the
    full [n, max_length] int64 result in the original row order, other ranks return None.
    Without an initialised process group this is the single-GPU loop."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return generate_batches(generate_fn, embeddings, max_length, batch_size, eos_token_id, in_flight)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(embeddings.shape[0], rank, world)
    mine = generate_batches(generate_fn, embeddings[lo:hi], max_length, batch_size, eos_token_id, in_flight)
    parts = [None] * world if rank == dst else None
    dist.gather_object((lo, hi, mine.numpy()), parts, dst=dst, group=group)  # host gather of token ids only
    if rank != dst:
        return None
    out = torch.full((embeddings.shape[0], max_length), eos_token_id, dtype=torch.int64)
    for plo, phi, arr in parts:
        out[plo:phi] = torch.from_numpy(arr)
    return out
