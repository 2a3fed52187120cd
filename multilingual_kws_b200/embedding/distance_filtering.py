"""The "embedding-extractor call" of the reference (multilingual_kws/embedding/distance_filtering.py:12-27)."""
from __future__ import annotations

import os
from pathlib import Path

from ..model import EmbeddingModel


def embedding_model(base_model_path=Path.home() / "tinyspeech_harvard/multilingual_embedding_wc/models/multilingual_context_73_0.8011",
                    base_model_output="dense_2", **kw) -> EmbeddingModel:
    """Loads the base classifier's weights and cuts the network at `base_model_output` (Keras layer name).
    Returns an object with .predict(specs[N,49,40(,1)]) -> [N,1024] and .trainable == False."""
    embedding = EmbeddingModel.load(os.fspath(base_model_path), output_layer=base_model_output, **kw)
    embedding.trainable = False
    return embedding
