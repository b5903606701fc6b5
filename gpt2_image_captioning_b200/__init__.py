"""gpt2_image_captioning_b200 -- the caption-generation hot path of thenoobychocobo/gpt2-image-captioning on B200.

Public API mirrors the reference's `src/models.py` (same class names / signatures); `generate` runs hand-written
sm_100a CUDA through libgic_b200.so (C ABI in include/gic_b200.h).  No Triton, no torch.compile, no CPU fallback.
"""
from .models import (ImageCaptioningModel, MLPMappingNetwork, RetrievalAggregator, RetrievalAugmentedTransformer,
                     TransformerMappingNetwork, accelerate)
from .engine import CaptionEngine
from .database import GpuFlatStore
from .sharding import shard_range, generate_sharded
from .inflight import map_batches
from .evaluation import generate_predictions, generate_and_evaluate, load_embeddings_pt, generate_for_embeddings

__all__ = ["ImageCaptioningModel", "MLPMappingNetwork", "TransformerMappingNetwork", "RetrievalAggregator",
           "RetrievalAugmentedTransformer", "accelerate", "CaptionEngine", "GpuFlatStore", "shard_range", "generate_sharded",
           "map_batches", "generate_predictions", "generate_and_evaluate", "load_embeddings_pt", "generate_for_embeddings"]
