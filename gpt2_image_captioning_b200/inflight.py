"""Several batches in flight on one GPU.

A decode step of one 1024-row batch is a chain of ~60 short kernels (6-17 us each, one persistent CTA per SM); the tail of
every kernel, the ~1 us hand-over to its successor and the HBM-bound attention leave SMs / the tensor pipe idle.  A second
batch on its own stream (own engine handle = own workspace, KV cache and CUDA graph; own host thread because the decode driver
polls the early-exit flag) fills those holes: measured 41.2k -> 45.4k captions/s on one B200 (tools/experiments/two_in_flight.py), token ids
identical.  Rows of different batches never interact, so this is the same job in a different order.

`map_batches(fn, batches, in_flight)` runs `fn(batch)` for every batch, batch i on worker i % in_flight.  Each worker has a
CUDA stream that first waits for the caller's current stream and that the caller's stream waits for at the end, so the call
behaves like the sequential loop for whoever consumes the results (and can be bracketed by CUDA events on the caller's stream).
Inside a worker `current_slot()` names its engine slot; `_EngineMixin._get_engine()` (models.py) keeps one engine CONTEXT per slot
(slot 0 owns the packed weights, the others are contexts cloned from it: own stream, workspace, KV cache and CUDA graphs).

Stream-ordering contract of `fn`'s results: a CUDA tensor returned by `fn` was allocated and written on the slot's stream.  The
caller's stream waits for every slot stream before `map_batches` returns, and the results are handed to the caching allocator's
bookkeeping with `record_stream(caller)`, so consuming or freeing them on the caller's stream afterwards is safe.  (Host tensors --
what `generate_batches` and `evaluation.run` return -- need none of this.)
"""
from __future__ import annotations

import threading
from typing import Callable, Sequence

import torch

_tls = threading.local()


def current_slot() -> int:
    return getattr(_tls, "slot", 0)


_streams: dict[tuple[int, int], "torch.cuda.Stream"] = {}


def _slot_stream(device: int, slot: int):
    """One long-lived stream per (device, slot): the caching allocator keeps a pool per stream, so fresh streams on every call would
    mean fresh cudaMallocs for every call's buffers."""
    key = (device, slot)
    st = _streams.get(key)
    if st is None:
        st = _streams[key] = torch.cuda.Stream(device)
    return st


def map_batches(fn: Callable, batches: Sequence, in_flight: int = 2) -> list:
    n = len(batches)
    in_flight = max(1, min(int(in_flight), n))
    if in_flight == 1:
        return [fn(b) for b in batches]
    cuda = torch.cuda.is_available()
    out: list = [None] * n
    errors: list = []
    device = torch.cuda.current_device() if cuda else None
    caller = torch.cuda.current_stream() if cuda else None
    streams = [_slot_stream(device, j) for j in range(in_flight)] if cuda else [None] * in_flight

    def work(slot: int) -> None:
        _tls.slot = slot
        try:
            if cuda:
                torch.cuda.set_device(device)
                streams[slot].wait_stream(caller)
                with torch.cuda.stream(streams[slot]):
                    for i in range(slot, n, in_flight):
                        out[i] = fn(batches[i])
            else:
                for i in range(slot, n, in_flight):
                    out[i] = fn(batches[i])
        except BaseException as exc:  # re-raised in the caller
            errors.append(exc)
        finally:
            _tls.slot = 0

    threads = [threading.Thread(target=work, args=(j,), name=f"gic-inflight-{j}") for j in range(in_flight)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if cuda:
        for s in streams:
            caller.wait_stream(s)
        for r in out:  # results allocated on a slot stream are about to be used (and freed) on the caller's
            for t in (r if isinstance(r, (tuple, list)) else (r,)):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(caller)
    if errors:
        raise errors[0]
    return out
