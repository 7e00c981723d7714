"""ORACLE — test infrastructure only.  CPU restatement of the encoder hand-off (pooling +
L2 normalisation of the bi-encoder's hidden states).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.

Parity status: **pinned** for last-token pooling + normalisation — ``tests/golden/encode.npz``
holds the outputs of the reference's own ``last_token_pool``
(``Latest/cuVS-2-gpu/old/generate_embeddings.py:11-21``) followed by its
``F.normalize(embeddings, p=2, dim=1)`` (``:103``), produced by importing the reference
(``tests/golden/make_golden_encode.py``); ``tests/test_oracle.py`` checks this file against them.
Mean pooling is the sentence-transformers ``Pooling`` module's masked mean (the reference calls
``SentenceTransformer.encode``, ``prepare_dataset.py:149``; un-vendored wheel, restated from its
published definition) and is **unpinned**.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def last_token_pool(hidden: np.ndarray, mask: Optional[np.ndarray]) -> np.ndarray:
    """generate_embeddings.py:11-21.  hidden [B,T,D], mask [B,T] integer (None = all ones)."""
    b, t, _ = hidden.shape
    if mask is None or int(mask[:, -1].sum()) == b:       # :15 left_padding
        return hidden[:, -1].copy()
    lengths = mask.sum(axis=1).astype(np.int64) - 1        # :19; -1 indexes the last token
    return hidden[np.arange(b), lengths].copy()


def mean_pool(hidden: np.ndarray, mask: Optional[np.ndarray]) -> np.ndarray:
    """sentence-transformers Pooling(mean): sum(h * mask) / clamp(sum(mask), min=1e-9), fp32."""
    h = hidden.astype(np.float32)
    if mask is None:
        return h.mean(axis=1, dtype=np.float32)
    m = mask.astype(np.float32)[:, :, None]
    return (h * m).sum(axis=1, dtype=np.float32) / np.maximum(m.sum(axis=1), np.float32(1e-9))


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(x, p=2, dim=1): x / max(||x||_2, eps) (:103)."""
    x = x.astype(np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
    return x / np.maximum(n, np.float32(eps))


def pool_normalize(hidden: np.ndarray, mask: Optional[np.ndarray], pooling: str = "last_token",
                   normalize: bool = True) -> np.ndarray:
    pooled = (last_token_pool(hidden, mask) if pooling == "last_token" else mean_pool(hidden, mask))
    pooled = pooled.astype(np.float32)
    return l2_normalize(pooled) if normalize else pooled
