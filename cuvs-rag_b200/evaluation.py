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
    """Same method names as the reference's RecallEvaluator (improved_multi_gpu_rag.py:310-357)."""

    @staticmethod
    def calculate_recall_at_k(retrieved_indices, relevant_indices, k: int) -> float:
        return recall_at_k(retrieved_indices, relevant_indices, k)

    @staticmethod
    def evaluate_recall_multiple_k(retrieved_indices, relevant_indices,
                                   k_values: List[int]) -> Dict[int, float]:
        return {k: recall_at_k(retrieved_indices, relevant_indices, k) for k in k_values}

    @staticmethod
    def batch_recall(retrieved: np.ndarray, truth: np.ndarray, k: int) -> float:
        """Mean recall@k over queries; ``truth`` [Q, >=k] holds each query's exact neighbours."""
        retrieved = np.asarray(retrieved)[:, :k]
        truth = np.asarray(truth)[:, :k]
        hits = 0
        for r, t in zip(retrieved, truth):
            hits += len(set(r.tolist()) & set(t.tolist()))
        return hits / float(truth.shape[0] * truth.shape[1])

    @staticmethod
    def generate_synthetic_ground_truth(num_queries: int, num_documents: int,
                                        relevance_ratio: float = 0.01, seed: int = 42):
        rng = np.random.default_rng(seed)
        per_query = max(1, int(num_documents * relevance_ratio))
        return [rng.choice(num_documents, size=per_query, replace=False).tolist()
                for _ in range(num_queries)]
