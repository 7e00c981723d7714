"""Recall metric used by the parity gate.

Formula restated from the reference (``improved_multi_gpu_rag.py:314-327`` and
``cuvs-2gpu-main.ipynb:L926-939``): ``recall@k = |top_k(retrieved) ∩ relevant| / |relevant|``.
The reference evaluates it against topic labels or random "ground truth" (always ~0, SURVEY.md
§3.6 bug 5); here ``relevant`` is the exact top-k of the same corpus, which makes it the usual
ANN recall.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np


def recall_at_k(retrieved: Sequence[int], relevant: Iterable[int], k: int) -> float:
    relevant = set(int(r) for r in relevant)
    if not relevant:
        return 0.0
    top = set(int(r) for r in list(retrieved)[:k])
    return len(top & relevant) / len(relevant)


class RecallEvaluator:
    """THE recall evaluator of the package (``improved_multi_gpu_rag`` re-exports this class):
    method names, signatures and corner cases exactly as the reference defines them
    (``improved_multi_gpu_rag.py:310-357``; pinned by ``tests/golden/recall.json``, generated from
    the imported reference), plus ``batch_recall`` for [Q, k] id matrices."""

    @staticmethod
    def calculate_recall_at_k(retrieved: np.ndarray, relevant: np.ndarray, k: int) -> float:
        retrieved, relevant = np.asarray(retrieved), np.asarray(relevant)
        if len(relevant) == 0:
            return 1.0 if len(retrieved) == 0 else 0.0
        top_k = retrieved[:k] if len(retrieved) >= k else retrieved
        return len(np.intersect1d(top_k, relevant)) / len(relevant)

    @staticmethod
    def evaluate_recall_multiple_k(retrieved: np.ndarray, relevant: np.ndarray,
                                   k_values: List[int]) -> Dict[int, float]:
        return {k: RecallEvaluator.calculate_recall_at_k(retrieved, relevant, min(k, len(retrieved)))
                for k in k_values}

    @staticmethod
    def generate_synthetic_ground_truth(num_queries: int, index_size: int,
                                        relevant_per_query: int = 100) -> Dict[int, np.ndarray]:
        np.random.seed(42)
        return {i: np.random.choice(index_size, size=min(relevant_per_query, index_size), replace=False)
                for i in range(num_queries)}

    @staticmethod
    def batch_recall(retrieved: np.ndarray, truth: np.ndarray, k: int) -> float:
        """Mean recall@k over queries; ``truth`` [Q, >=k] holds each query's exact neighbours."""
        retrieved = np.asarray(retrieved)[:, :k]
        truth = np.asarray(truth)[:, :k]
        hits = 0
        for r, t in zip(retrieved, truth):
            hits += len(set(r.tolist()) & set(t.tolist()))
        return hits / float(truth.shape[0] * truth.shape[1])
