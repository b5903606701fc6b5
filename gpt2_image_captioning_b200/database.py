"""GpuFlatStore: the retrieval database of the RAT variant resident in HBM.

Duck-types the reference's `FAISSStore` (src/database/faiss_store.py:16-52): `image_index.search(q, k)`,
`caption_index.reconstruct(i)`, `image_metadata`, `caption_metadata`, `filename_to_caption_indices`, `close()` -- so the
reference's own `retrieve_images_by_vector_similarity` / `get_caption_embeddings` can be driven through it -- and adds
the fused device path `retrieve_and_aggregate` used by `RetrievalAugmentedTransformer.generate`:
  exact inner-product top-(top_i+10) over the IMAGE matrix -> hit filter (idx != -1, score <= 0.9999) -> caption rows of
  the first top_i hits -> first top_k rows of the CAPTION matrix -> aggregate (zero padding rows count) -> query + pooled
(src/database/faiss_store.py:132-251, src/models.py:589-625,655-695), with no GPU->CPU->GPU hop.
Search is EXACT (IndexFlatIP semantics, ties -> lowest index): a quality superset of the reference's default HNSW index.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _capi, faiss_files, ops
from .inflight import current_slot


class _GpuFlatIndex:
    """faiss.IndexFlatIP-like view of one device matrix."""

    def __init__(self, matrix: torch.Tensor, tensor_cores: bool | None = None):
        self.matrix = matrix
        self.ntotal, self.d = int(matrix.shape[0]), int(matrix.shape[1])
        self._ws: dict[int, torch.Tensor] = {}  # scan workspace per in-flight slot (inflight.py): slots search concurrently
        # tensor-core scan (bf16x2 tcgen05 candidates + exact re-scoring + certificate, include/gic_b200.h gic_topk_ip_tc): same
        # scores and indices as the exact fp32 scan.  GIC_RETRIEVAL_EXACT=1 keeps the fp32 CUDA-core scan.
        if tensor_cores is None:
            tensor_cores = os.environ.get("GIC_RETRIEVAL_EXACT", "0") != "1"
        self.hi = self.lo = None
        self.norm_max = 0.0
        if tensor_cores and self.ntotal > 0 and self.d % 64 == 0 and self.d <= 2048 and matrix.is_cuda:
            self.hi = torch.empty(matrix.shape, dtype=torch.bfloat16, device=matrix.device)
            self.lo = torch.empty(matrix.shape, dtype=torch.bfloat16, device=matrix.device)
            with torch.cuda.device(matrix.device):
                ops.pack_bf16x2(matrix, self.hi, self.lo)
            self.norm_max = float(torch.linalg.vector_norm(matrix, dim=1).max().item())

    def search_device(self, q: torch.Tensor, k: int):
        q = q.to(device=self.matrix.device, dtype=torch.float32).contiguous()
        if q.dim() == 1:
            q = q.unsqueeze(0)
        B = q.shape[0]
        scores = torch.empty(B, k, dtype=torch.float32, device=q.device)
        idx = torch.empty(B, k, dtype=torch.int64, device=q.device)
        if B == 0:
            return scores, idx
        lib = _capi.lib()
        tc = self.hi is not None and bool(lib.gic_topk_tc_supported(self.d, k))
        need = int((lib.gic_topk_tc_workspace_bytes if tc else lib.gic_topk_workspace_bytes)(B, self.ntotal, self.d, k))
        slot = current_slot()
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < need:
            self._ws.pop(slot, None)
            ws = self._ws[slot] = torch.empty(need, dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            if tc:
                ops.topk_ip_tc(q, self.matrix, self.hi, self.lo, self.norm_max, k, scores, idx, ws)
            else:
                ops.topk_ip(q, self.matrix, k, scores, idx, ws)
        return scores, idx

    def search(self, x, k: int):
        """(float32 [B,k] descending, int64 [B,k]; -1 / -inf padding when k > ntotal) as numpy, like faiss."""
        q = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        s, i = self.search_device(q, k)
        return s.cpu().numpy(), i.cpu().numpy()

    def reconstruct(self, i: int) -> np.ndarray:
        return self.matrix[int(i)].cpu().numpy()


class GpuFlatStore:
    def __init__(self, image_embeddings, caption_embeddings, image_metadata: list[str], caption_metadata: list[dict],
                 device: torch.device | str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("GpuFlatStore needs a CUDA device; there is no CPU fallback")
        dev = torch.device(device)
        img = torch.as_tensor(image_embeddings, dtype=torch.float32).to(dev).contiguous()
        cap = torch.as_tensor(caption_embeddings, dtype=torch.float32).to(dev).contiguous()
        if img.dim() != 2 or cap.dim() != 2 or img.shape[1] != cap.shape[1]:
            raise ValueError("image and caption matrices must be [N, D] with the same D")
        if len(image_metadata) != img.shape[0] or len(caption_metadata) != cap.shape[0]:
            raise ValueError("metadata lengths must match the matrices")
        self.image_index, self.caption_index = _GpuFlatIndex(img), _GpuFlatIndex(cap)
        self.image_metadata, self.caption_metadata = image_metadata, caption_metadata
        self.filename_to_caption_indices: dict[str, list[int]] = {}
        for row, meta in enumerate(caption_metadata):  # same reverse lookup as faiss_store.py:42-48
            self.filename_to_caption_indices.setdefault(meta["filename"], []).append(row)
        # CSR over images: caption rows of image i = cap_row_ids[start[i]:start[i+1]]
        counts = [len(self.filename_to_caption_indices.get(f, ())) for f in image_metadata]
        start = np.zeros(len(image_metadata) + 1, np.int64)
        np.cumsum(counts, out=start[1:])
        flat = np.fromiter((r for f in image_metadata for r in self.filename_to_caption_indices.get(f, ())), np.int64,
                           count=int(start[-1]))
        self.cap_row_start = torch.from_numpy(start).to(dev)
        self.cap_row_ids = torch.from_numpy(flat).to(dev) if len(flat) else torch.zeros(1, dtype=torch.int64, device=dev)
        self.device = dev

    @classmethod
    def from_faiss_store(cls, store, device="cuda") -> "GpuFlatStore":
        """Upload an existing reference FAISSStore (flat or HNSW-flat: both keep the raw vectors) to HBM."""
        def matrix(index):
            if hasattr(index, "reconstruct_n"):
                return np.asarray(index.reconstruct_n(0, index.ntotal), dtype=np.float32)
            return np.stack([index.reconstruct(i) for i in range(index.ntotal)]).astype(np.float32)
        return cls(matrix(store.image_index), matrix(store.caption_index), store.image_metadata, store.caption_metadata, device)

    @classmethod
    def from_directory(cls, db_directory: str, device="cuda") -> "GpuFlatStore | None":
        """`create_faiss_store(db_directory)` (src/database/faiss_store.py:55-104) without faiss: reads the raw vectors out
        of `image_index.faiss` / `caption_index.faiss` (flat or HNSW-flat, faiss_files.py) plus the two metadata pickles
        and uploads them; None when the directory is incomplete, like the reference."""
        loaded = faiss_files.read_store_directory(db_directory)
        if loaded is None:
            return None
        return cls(*loaded, device=device)

    @classmethod
    def from_embedding_files(cls, image_embedding_file_path: str, caption_embedding_file_path: str, device="cuda") -> "GpuFlatStore":
        """`run_faiss_indexing_pipeline` (src/database/faiss_indexing.py:18-150) for the exact GPU store: the "index build"
        is the upload of the two matrices.  Image file: `{"filenames", "embeddings"}`; caption file: list of
        `{"filenames": name, "embeddings": [{"embedding", "caption_id"}, ...]}` (captions of unknown images are skipped)."""
        image_data = torch.load(image_embedding_file_path, weights_only=True)
        caption_data = torch.load(caption_embedding_file_path, weights_only=False)
        names = list(image_data["filenames"])
        img = image_data["embeddings"].to(torch.float32)
        cap, meta = faiss_files.flatten_caption_entries(caption_data, names)
        if cap.size == 0:
            cap = np.zeros((0, img.shape[1]), np.float32)
        return cls(img, cap, names, meta, device)

    def save(self, db_directory: str) -> None:
        """`save_faiss_store` (src/database/faiss_store.py:107-129): the same four files, indices written as IndexFlatIP,
        readable by the reference's `create_faiss_store` and by `from_directory`."""
        faiss_files.write_store_directory(db_directory, self.image_index.matrix.cpu().numpy(), self.caption_index.matrix.cpu().numpy(),
                                          self.image_metadata, self.caption_metadata)

    def close(self) -> None:  # API compatibility (faiss_store.py:50-52)
        pass

    # ---- fused device path ----------------------------------------------------------------------------------------------
    def retrieve_rows(self, image_embeddings: torch.Tensor, top_i: int, top_k: int) -> torch.Tensor:
        """int64 [B, top_k] caption-matrix rows (-1 = zero padding)."""
        scores, idx = self.image_index.search_device(image_embeddings, top_i + 10)  # faiss_store.py:153-155
        rows = torch.empty(scores.shape[0], top_k, dtype=torch.int64, device=self.device)
        if scores.shape[0]:
            with torch.cuda.device(self.device):
                ops.select_caption_rows(scores, idx, self.cap_row_start, self.cap_row_ids, top_i, top_k, rows)
        return rows

    def retrieve_caption_embeddings(self, image_embeddings: torch.Tensor, top_i: int, top_k: int) -> torch.Tensor:
        """float32 [B, top_k, D], zero rows for padding -- what `_retrieve_batch` returns (src/models.py:655-695)."""
        rows = self.retrieve_rows(image_embeddings, top_i, top_k)
        out = torch.empty(rows.shape[0], top_k, self.caption_index.d, dtype=torch.float32, device=self.device)
        if rows.shape[0]:
            with torch.cuda.device(self.device):
                ops.gather_caption_rows(self.caption_index.matrix, rows, out)
        return out

    def retrieve_and_aggregate(self, image_embeddings: torch.Tensor, top_i: int, top_k: int, aggregation: str = "mean",
                               attention_weight: torch.Tensor | None = None, attention_bias: torch.Tensor | None = None) -> torch.Tensor:
        """query + pooled retrieved captions, float32 [B, D] on the input's device (src/models.py:763-768).  aggregation
        "attention" takes RetrievalAggregator.attention_proj's weight [1, D] and bias [1] (:606-616)."""
        if aggregation not in _capi.AGG and aggregation != "attention":
            raise ValueError(f"Unknown aggregation_type: {aggregation}")
        if aggregation == "attention":
            if attention_weight is None or attention_bias is None:
                raise ValueError("attention pooling needs attention_proj's weight and bias")
            src_device = image_embeddings.device
            q = image_embeddings.to(device=self.device, dtype=torch.float32).contiguous()
            rows = self.retrieve_rows(q, top_i, top_k)
            out = torch.empty_like(q)
            if q.shape[0]:
                w = attention_weight.detach().to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
                bb = attention_bias.detach().to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
                with torch.cuda.device(self.device):
                    ops.gather_attention_add(q, self.caption_index.matrix, rows, w, bb, out)
            return out.to(src_device)
        src_device = image_embeddings.device
        q = image_embeddings.to(device=self.device, dtype=torch.float32).contiguous()
        rows = self.retrieve_rows(q, top_i, top_k)
        out = torch.empty_like(q)
        if q.shape[0]:
            with torch.cuda.device(self.device):
                ops.gather_aggregate_add(q, self.caption_index.matrix, rows, _capi.AGG[aggregation], out)
        return out.to(src_device)
