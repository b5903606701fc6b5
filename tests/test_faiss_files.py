"""CPU checks of the faiss_db/ directory reader (gpt2_image_captioning_b200/faiss_files.py).
faiss is not installable here, so the files are produced by `write_flat_ip` and by the byte-level HNSW writer below, both
following the layout stated in faiss_files.py's header (restated from faiss's index_write.cpp; unpinned by a faiss-written file)."""
import os
import pickle
import struct

import numpy as np
import pytest
import torch

from gpt2_image_captioning_b200 import faiss_files as ff


def _vec(fmt: str, values) -> bytes:
    values = list(values)
    return struct.pack("<Q", len(values)) + struct.pack(f"<{len(values)}{fmt}", *values)


def _write_hnsw_flat(path, matrix, trailer_ints=5, M=4, metric=0, nested=None):
    """IndexHNSWFlat as faiss serialises it: "IHNf", header, graph vectors, scalars, nested flat index."""
    n, d = matrix.shape
    rng = np.random.default_rng(0)
    levels = [1] * n
    offsets = [2 * M * i for i in range(n + 1)]
    neighbors = rng.integers(-1, max(n, 1), 2 * M * n).tolist()
    with open(path, "wb") as f:
        f.write(b"IHNf")
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric))
        f.write(_vec("d", [0.9, 0.1]))               # assign_probas
        f.write(_vec("i", [0, 2 * M, 3 * M]))       # cum_nneighbor_per_level
        f.write(_vec("i", levels))
        f.write(_vec("Q", offsets))
        f.write(_vec("i", neighbors))
        f.write(struct.pack(f"<{trailer_ints}i", *([0, 0, 200, 64, 1][:trailer_ints])))
    tmp = str(path) + ".flat"
    ff.write_flat_ip(tmp, matrix if nested is None else nested)
    with open(path, "ab") as f, open(tmp, "rb") as g:
        f.write(g.read())
    os.remove(tmp)


def _matrix(n, d, seed=0):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def test_flat_round_trip_and_layout(tmp_path):
    m = _matrix(37, 16)
    p = tmp_path / "i.faiss"
    ff.write_flat_ip(p, m)
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxFI" and len(raw) == 4 + 33 + 8 + m.nbytes
    assert struct.unpack_from("<iq", raw, 4) == (16, 37) and struct.unpack_from("<Q", raw, 37) == (37 * 16,)
    assert ff.locate_vectors(p) == (16, 37, 0, 45, "flat")
    assert np.array_equal(ff.read_vectors(p), m)
    mm = ff.read_vectors(p, mmap=True)
    assert isinstance(mm, np.memmap) and np.array_equal(mm, m)


def test_flat_l2_and_empty(tmp_path):
    m = _matrix(5, 8)
    p = tmp_path / "l2.faiss"
    ff.write_flat_ip(p, m)
    raw = bytearray(open(p, "rb").read())
    raw[:4] = b"IxF2"
    struct.pack_into("<i", raw, 4 + 29, 1)
    open(p, "wb").write(raw)
    assert ff.locate_vectors(p)[2] == 1 and np.array_equal(ff.read_vectors(p), m)
    ff.write_flat_ip(p, np.zeros((0, 8), np.float32))
    assert ff.read_vectors(p).shape == (0, 8)


@pytest.mark.parametrize("trailer_ints", [5, 4])
def test_hnsw_flat_vectors_are_found_behind_the_graph(tmp_path, trailer_ints):
    m = _matrix(50, 32, 3)
    p = tmp_path / "h.faiss"
    _write_hnsw_flat(p, m, trailer_ints)
    d, n, metric, off, kind = ff.locate_vectors(p)
    assert (d, n, metric, kind) == (32, 50, 0, "hnsw_flat")
    assert off == os.path.getsize(p) - m.nbytes
    assert np.array_equal(ff.read_vectors(p), m)


def test_malformed_files_raise(tmp_path):
    m = _matrix(20, 8)
    p = tmp_path / "x.faiss"
    ff.write_flat_ip(p, m)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-4])                                 # truncated vectors
    with pytest.raises(ff.FaissFormatError, match="ends inside"):
        ff.read_vectors(p)
    open(p, "wb").write(raw[:20])                                 # truncated header
    with pytest.raises(ff.FaissFormatError, match="truncated"):
        ff.read_vectors(p)
    open(p, "wb").write(b"IwFl" + raw[4:])                        # IVF index: no raw vectors in a flat layout
    with pytest.raises(ff.FaissFormatError, match="unsupported index type"):
        ff.read_vectors(p)
    bad = bytearray(raw)
    struct.pack_into("<Q", bad, 37, 20 * 8 + 1)                   # count != ntotal * d
    open(p, "wb").write(bad)
    with pytest.raises(ff.FaissFormatError, match="vector count"):
        ff.read_vectors(p)
    open(p, "wb").write(b"")
    with pytest.raises(ff.FaissFormatError):
        ff.read_vectors(p)
    _write_hnsw_flat(p, m, nested=_matrix(19, 8))                # nested storage disagrees with the HNSW header
    with pytest.raises(ff.FaissFormatError, match="nested storage"):
        ff.read_vectors(p)
    _write_hnsw_flat(p, m, trailer_ints=3)                        # storage not where any release puts it
    with pytest.raises(ff.FaissFormatError, match="no flat storage"):
        ff.read_vectors(p)
    with pytest.raises(ValueError):
        ff.write_flat_ip(p, np.zeros(8, np.float32))


def test_store_directory_round_trip_and_missing_files(tmp_path):
    img, cap = _matrix(6, 8, 1), _matrix(13, 8, 2)
    names = [f"img_{i}.jpg" for i in range(6)]
    meta = [{"filename": names[j % 6], "caption_id": 100 + j} for j in range(13)]
    d = tmp_path / "faiss_db"
    assert ff.read_store_directory(d) is None                      # create_faiss_store: nothing there -> None
    ff.write_store_directory(d, img, cap, names, meta)
    assert sorted(os.listdir(d)) == ["caption_index.faiss", "caption_metadata.pkl", "image_index.faiss", "image_metadata.pkl"]
    gi, gc, gn, gm = ff.read_store_directory(d)
    assert np.array_equal(gi, img) and np.array_equal(gc, cap) and gn == names and gm == meta
    _write_hnsw_flat(d / "image_index.faiss", img)                  # the reference's default: HNSW image + caption indices
    _write_hnsw_flat(d / "caption_index.faiss", cap)
    gi, gc, _, _ = ff.read_store_directory(d)
    assert np.array_equal(gi, img) and np.array_equal(gc, cap)
    with open(d / "image_metadata.pkl", "wb") as f:
        pickle.dump(names[:-1], f)
    with pytest.raises(ff.FaissFormatError, match="metadata lengths"):
        ff.read_store_directory(d)
    os.remove(d / "caption_metadata.pkl")
    assert ff.read_store_directory(d) is None


def test_flatten_caption_entries_follows_the_indexing_pipeline():
    """src/database/faiss_indexing.py:84-116 restated: file order, unknown images skipped, tensors and lists accepted."""
    names = ["a.jpg", "b.jpg"]
    data = [
        {"filenames": "b.jpg", "embeddings": [{"embedding": torch.tensor([1.0, 2.0]), "caption_id": 7},
                                              {"embedding": [3.0, 4.0], "caption_id": 8}]},
        {"filenames": "zzz.jpg", "embeddings": [{"embedding": torch.tensor([9.0, 9.0]), "caption_id": 1}]},
        {"filenames": "a.jpg", "embeddings": [{"embedding": np.array([5.0, 6.0]), "caption_id": 9}]},
    ]
    m, meta = ff.flatten_caption_entries(data, names)
    assert m.dtype == np.float32 and m.tolist() == [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]
    assert meta == [{"filename": "b.jpg", "caption_id": 7}, {"filename": "b.jpg", "caption_id": 8}, {"filename": "a.jpg", "caption_id": 9}]
    m, meta = ff.flatten_caption_entries([], names)
    assert m.size == 0 and meta == []


class _NumpyFlatIP:
    """Stand-in for faiss.IndexFlatIP / IndexHNSWFlat with exactly what the reference's pipeline touches."""

    def __init__(self, d, *_):
        self.d, self.rows, self.hnsw = d, np.zeros((0, d), np.float32), type("H", (), {})()

    def add(self, x):
        assert x.dtype == np.float32 and x.shape[1] == self.d
        self.rows = np.concatenate([self.rows, x])

    @property
    def ntotal(self):
        return self.rows.shape[0]


def test_directory_written_by_the_reference_pipeline(tmp_path, monkeypatch):
    """The unmodified run_faiss_indexing_pipeline + save_faiss_store (src/database/faiss_indexing.py:18-150,
    faiss_store.py:107-129), with numpy-backed index objects in place of faiss and write_index -> write_flat_ip, must
    leave a directory that read_store_directory / flatten_caption_entries reproduce (file names, row order, metadata)."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    ref_harness.import_reference()
    import importlib
    import sys
    fake = sys.modules["faiss"]
    monkeypatch.setattr(fake, "IndexFlatIP", _NumpyFlatIP, raising=False)
    monkeypatch.setattr(fake, "IndexHNSWFlat", _NumpyFlatIP, raising=False)
    monkeypatch.setattr(fake, "METRIC_INNER_PRODUCT", 0, raising=False)
    monkeypatch.setattr(fake, "write_index", lambda index, path: ff.write_flat_ip(path, index.rows), raising=False)
    pipeline = importlib.import_module("src.database.faiss_indexing")

    g = torch.Generator().manual_seed(5)
    names = [f"{i:012d}.jpg" for i in range(9)]
    img = torch.nn.functional.normalize(torch.randn(9, 16, generator=g), dim=1)
    caps = []
    for j, name in enumerate(names[::-1] + ["not_in_the_image_file.jpg"]):   # caption file order != image order
        caps.append({"filenames": name, "embeddings": [{"embedding": torch.randn(16, generator=g), "caption_id": 10 * j + c}
                                                        for c in range(j % 4)]})
    torch.save({"filenames": names, "embeddings": img}, tmp_path / "img.pt")
    torch.save(caps, tmp_path / "cap.pt")
    for approximate in (True, False):
        d = tmp_path / f"db_{approximate}"
        pipeline.run_faiss_indexing_pipeline(str(d), str(tmp_path / "img.pt"), str(tmp_path / "cap.pt"), use_approximate=approximate)
        gi, gc, gn, gm = ff.read_store_directory(d)
        want_cap, want_meta = ff.flatten_caption_entries(caps, names)
        assert np.array_equal(gi, img.numpy()) and gn == names
        assert np.array_equal(gc, want_cap) and gm == want_meta
        assert len(gm) == sum(j % 4 for j in range(9)) and all(m["filename"] in names for m in gm)
