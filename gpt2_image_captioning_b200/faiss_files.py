"""Reader / writer for the reference's on-disk retrieval store (`faiss_db/`), without faiss.

The reference keeps its RAT database as four files (src/database/faiss_store.py:78-104,107-129):
  image_index.faiss, caption_index.faiss   -- `faiss.write_index` of an IndexHNSWFlat (default, src/database/faiss_indexing.py:63-71)
                                              or an IndexFlatIP (`use_approximate=False`, :72-75)
  image_metadata.pkl, caption_metadata.pkl -- pickled list[str] / list[{"filename", "caption_id"}]
Both index kinds keep the raw fp32 vectors (HNSW-*Flat* stores them in a nested flat index), which is all the exact GPU
store needs: this module pulls the `[ntotal, d]` matrix out of the file and `GpuFlatStore.from_directory` uploads it.

faiss-cpu 1.13.1 (uv.lock:602-603) is not installable here, so the layout below is restated from faiss's published
serialisation (faiss/impl/index_write.cpp, index_read.cpp) and is **unpinned by a faiss-written file**; the reader
therefore validates everything it can (fourcc, d, ntotal, payload length against the file size) and raises on anything else:

  flat index   : fourcc "IxFI" (inner product) | "IxF2" (L2)                       4 bytes
                 header: d int32, ntotal int64, dummy int64 x2, is_trained u8, metric_type int32 [, metric_arg f32 if metric_type > 1]
                 vectors: count uint64 (= ntotal * d floats), then count fp32, row-major
  HNSW-flat    : fourcc "IHNf", the same header, then the graph:
                 assign_probas (u64 n, n f64), cum_nneighbor_per_level (u64 n, n i32), levels (u64 n, n i32),
                 offsets (u64 n, n u64), neighbors (u64 n, n i32), entry_point, max_level, efConstruction, efSearch,
                 upper_beam (int32 each; the last one is a dummy in current releases), then the nested flat index as above.
The graph is skipped, not interpreted: search on the GPU is exact (a quality superset of HNSW).
"""
from __future__ import annotations

import os
import pickle
import struct

import numpy as np

IMAGE_INDEX, CAPTION_INDEX = "image_index.faiss", "caption_index.faiss"
IMAGE_META, CAPTION_META = "image_metadata.pkl", "caption_metadata.pkl"
_FLAT = {b"IxFI": 0, b"IxF2": 1}  # fourcc -> faiss MetricType (METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1)
_HNSW_FLAT = b"IHNf"
_DUMMY = 1 << 20


class FaissFormatError(ValueError):
    pass


def _header(buf: memoryview, pos: int, path: str) -> tuple[int, int, int, int]:
    """(d, ntotal, metric_type, next offset) of the common index header starting at `pos` (just after the fourcc)."""
    if pos + 33 > len(buf):
        raise FaissFormatError(f"{path}: truncated index header")
    d, ntotal, _d0, _d1, _trained, metric = struct.unpack_from("<iqqqBi", buf, pos)
    pos += 33
    if metric > 1:  # metric_arg follows for the parametrised metrics
        pos += 4
    if d <= 0 or ntotal < 0 or metric < 0:
        raise FaissFormatError(f"{path}: implausible header (d={d}, ntotal={ntotal}, metric={metric})")
    return d, ntotal, metric, pos


def _flat_payload(buf: memoryview, pos: int, path: str) -> tuple[int, int, int, int]:
    """Parse a flat index whose fourcc starts at `pos`: (d, ntotal, metric, byte offset of the fp32 matrix)."""
    tag = bytes(buf[pos:pos + 4])
    if tag not in _FLAT:
        raise FaissFormatError(f"{path}: expected a flat index (IxFI / IxF2) at byte {pos}, found {tag!r}")
    d, ntotal, metric, pos = _header(buf, pos + 4, path)
    if pos + 8 > len(buf):
        raise FaissFormatError(f"{path}: truncated before the vector count")
    (count,) = struct.unpack_from("<Q", buf, pos)
    pos += 8
    if count != ntotal * d:
        raise FaissFormatError(f"{path}: vector count {count} != ntotal * d = {ntotal} * {d}")
    if pos + 4 * count > len(buf):
        raise FaissFormatError(f"{path}: file ends inside the vectors ({len(buf) - pos} of {4 * count} bytes)")
    return d, ntotal, metric, pos


def _skip_vector(buf: memoryview, pos: int, itemsize: int, path: str, what: str) -> tuple[int, int]:
    if pos + 8 > len(buf):
        raise FaissFormatError(f"{path}: truncated before {what}")
    (n,) = struct.unpack_from("<Q", buf, pos)
    end = pos + 8 + n * itemsize
    if end > len(buf):
        raise FaissFormatError(f"{path}: {what} ({n} items) runs past the end of the file")
    return n, end


def locate_vectors(path: str) -> tuple[int, int, int, int, str]:
    """(d, ntotal, metric_type, byte offset of the fp32 [ntotal, d] matrix, kind) for a flat or HNSW-flat index file."""
    with open(path, "rb") as f:
        mm = np.memmap(f, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
    buf = memoryview(mm)
    tag = bytes(buf[:4])
    if tag in _FLAT:
        d, ntotal, metric, off = _flat_payload(buf, 0, path)
        return d, ntotal, metric, off, "flat"
    if tag != _HNSW_FLAT:
        raise FaissFormatError(f"{path}: unsupported index type {tag!r}: the store needs the raw vectors, i.e. IndexFlatIP / "
                               f"IndexFlatL2 / IndexHNSWFlat (what src/database/faiss_indexing.py builds)")
    d, ntotal, metric, pos = _header(buf, 4, path)
    _, pos = _skip_vector(buf, pos, 8, path, "hnsw.assign_probas")
    _, pos = _skip_vector(buf, pos, 4, path, "hnsw.cum_nneighbor_per_level")
    n_levels, pos = _skip_vector(buf, pos, 4, path, "hnsw.levels")
    n_offsets, pos = _skip_vector(buf, pos, 8, path, "hnsw.offsets")
    _, pos = _skip_vector(buf, pos, 4, path, "hnsw.neighbors")
    if n_levels != ntotal or n_offsets != ntotal + 1:
        raise FaissFormatError(f"{path}: HNSW graph covers {n_levels} nodes / {n_offsets} offsets, header says ntotal = {ntotal}")
    # entry_point, max_level, efConstruction, efSearch [, upper_beam]: releases differ in whether the fifth int is written,
    # so accept the nested flat index after either four or five of them.
    for ints in (5, 4):
        at = pos + 4 * ints
        if bytes(buf[at:at + 4]) in _FLAT:
            sd, sn, smetric, off = _flat_payload(buf, at, path)
            if (sd, sn) != (d, ntotal):
                raise FaissFormatError(f"{path}: nested storage is [{sn}, {sd}], HNSW header says [{ntotal}, {d}]")
            return d, ntotal, smetric, off, "hnsw_flat"
    raise FaissFormatError(f"{path}: no flat storage index after the HNSW graph (byte {pos})")


def read_vectors(path: str, mmap: bool = False) -> np.ndarray:
    """fp32 [ntotal, d] matrix of a flat / HNSW-flat index file (`mmap=True`: read-only view of the file, no host copy
    before the upload)."""
    d, ntotal, _metric, off, _kind = locate_vectors(path)
    if ntotal == 0:
        return np.zeros((0, d), np.float32)
    if mmap:
        return np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(ntotal, d))
    return np.fromfile(path, dtype="<f4", count=ntotal * d, offset=off).reshape(ntotal, d)


def write_flat_ip(path: str, matrix: np.ndarray) -> None:
    """Write `matrix` as an IndexFlatIP file (what `faiss.write_index(faiss.IndexFlatIP(d))` produces after `add`), so a
    store built or filtered on the GPU can be handed back to the reference's `create_faiss_store`."""
    m = np.ascontiguousarray(matrix, dtype="<f4")
    if m.ndim != 2 or m.shape[1] <= 0:
        raise ValueError("matrix must be [ntotal, d] with d > 0")
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(struct.pack("<iqqqBi", m.shape[1], m.shape[0], _DUMMY, _DUMMY, 1, 0))
        f.write(struct.pack("<Q", m.size))
        m.tofile(f)


def read_store_directory(db_directory: str, mmap: bool = False):
    """(image matrix, caption matrix, image_metadata, caption_metadata) of a reference `faiss_db/` directory, or None when
    any of the four files is missing (create_faiss_store returns None then, src/database/faiss_store.py:83-93)."""
    paths = [os.path.join(db_directory, n) for n in (IMAGE_INDEX, CAPTION_INDEX, IMAGE_META, CAPTION_META)]
    if not all(os.path.exists(p) for p in paths):
        return None
    img, cap = read_vectors(paths[0], mmap), read_vectors(paths[1], mmap)
    with open(paths[2], "rb") as f:
        image_metadata = pickle.load(f)
    with open(paths[3], "rb") as f:
        caption_metadata = pickle.load(f)
    if len(image_metadata) != img.shape[0] or len(caption_metadata) != cap.shape[0]:
        raise FaissFormatError(f"{db_directory}: metadata lengths ({len(image_metadata)}, {len(caption_metadata)}) do not match "
                               f"the indices ({img.shape[0]}, {cap.shape[0]})")
    return img, cap, image_metadata, caption_metadata


def write_store_directory(db_directory: str, image_matrix, caption_matrix, image_metadata, caption_metadata) -> None:
    """Same four files as save_faiss_store (src/database/faiss_store.py:107-129), the indices as IndexFlatIP."""
    os.makedirs(db_directory, exist_ok=True)
    write_flat_ip(os.path.join(db_directory, IMAGE_INDEX), image_matrix)
    write_flat_ip(os.path.join(db_directory, CAPTION_INDEX), caption_matrix)
    with open(os.path.join(db_directory, IMAGE_META), "wb") as f:
        pickle.dump(list(image_metadata), f)
    with open(os.path.join(db_directory, CAPTION_META), "wb") as f:
        pickle.dump(list(caption_metadata), f)


def flatten_caption_entries(caption_data, image_filenames) -> tuple[np.ndarray, list[dict]]:
    """The caption half of run_faiss_indexing_pipeline (src/database/faiss_indexing.py:84-116): entries
    `{"filenames": name, "embeddings": [{"embedding": vec, "caption_id": id}, ...]}` -> (fp32 [n_captions, d], metadata),
    skipping images that are not in `image_filenames`, in file order."""
    known = set(image_filenames)
    rows, meta = [], []
    for entry in caption_data:
        name = entry["filenames"]
        if name not in known:
            continue
        for cap in entry["embeddings"]:
            e = cap["embedding"]
            rows.append(np.asarray(e.numpy() if hasattr(e, "numpy") else e, dtype=np.float32))
            meta.append({"filename": name, "caption_id": cap["caption_id"]})
    matrix = np.stack(rows).astype(np.float32) if rows else np.zeros((0, 0), np.float32)
    return matrix, meta
