"""Encoder hand-off: from the bi-encoder's hidden states to search results
without leaving the GPU.

The reference encodes, pools, normalises, copies the embeddings to the host
(``generate_embeddings.py:100-105``: ``last_token_pool`` -> ``F.normalize`` ->
``embeddings.cpu().numpy()``) and copies the query back to every GPU before the search
(``cuvs-2gpu-main.ipynb`` cell 16: ``.cpu()`` then ``.to(device)``).  Here the hidden states are
pooled and normalised by ``b2vs_pool_normalize`` (``csrc/encode.cu``) straight into the query
matrix, in the dtype the index holds, and handed to ``SearchResultAggregator`` as a device tensor:
the only device-to-host copy left is the final ``[Q, k]`` result.

There is no CPU path: CPU tensors are rejected (the oracle restatement lives in ``oracle/encode.py``
and is test infrastructure).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Mapping, Optional

import torch

import _native
from search_result_aggregator import AggregatedSearchResult, SearchConfig, SearchResultAggregator

POOLING_MODES = ("last_token", "mean")


def last_token_pool(last_hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Same name, arguments and result as the reference helper
    (``generate_embeddings.py:11-21``), computed on the GPU by ``b2vs_pool_normalize``."""
    return _native.pool_normalize(last_hidden_states, attention_mask, "last_token", normalize=False)


def embed_queries(hidden: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                  pooling: str = "last_token", normalize: bool = True,
                  dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Pooled, L2-normalised query rows ``[B, D]`` on the hidden states' GPU, in ``dtype``."""
    if pooling not in POOLING_MODES:
        raise ValueError(f"unknown pooling {pooling!r}; expected one of {POOLING_MODES}")
    return _native.pool_normalize(hidden, attention_mask, pooling, normalize, dtype)


def _hidden_states_of(outputs: Any) -> torch.Tensor:
    """Hugging Face model outputs (``.last_hidden_state`` / first tuple element) or a bare tensor."""
    if isinstance(outputs, torch.Tensor):
        return outputs
    if hasattr(outputs, "last_hidden_state"):
        return outputs.last_hidden_state
    if isinstance(outputs, Mapping) and "last_hidden_state" in outputs:
        return outputs["last_hidden_state"]
    if isinstance(outputs, (tuple, list)) and outputs and isinstance(outputs[0], torch.Tensor):
        return outputs[0]
    raise TypeError(f"cannot find hidden states in encoder output of type {type(outputs).__name__}")


class QueryEncoderHandoff:
    """Runs ``encoder(**inputs)`` and the distributed search back to back on the device.

    ``encoder`` is any callable returning hidden states ``[B, T, D]`` (a Hugging Face ``AutoModel``
    as in the reference, ``generate_embeddings.py:54-67``); ``query_dtype`` is the dtype the
    indexes were built in (default: the encoder's output dtype).
    """

    def __init__(self, aggregator: SearchResultAggregator, encoder: Callable[..., Any],
                 pooling: str = "last_token", normalize: bool = True,
                 query_dtype: Optional[torch.dtype] = None):
        if pooling not in POOLING_MODES:
            raise ValueError(f"unknown pooling {pooling!r}; expected one of {POOLING_MODES}")
        if not callable(encoder):
            raise TypeError("encoder must be callable")
        self.aggregator = aggregator
        self.encoder = encoder
        self.pooling = pooling
        self.normalize = normalize
        self.query_dtype = query_dtype

    def embed(self, encoder_inputs: Mapping[str, torch.Tensor]) -> torch.Tensor:
        if not isinstance(encoder_inputs, Mapping) or not encoder_inputs:
            raise ValueError("encoder_inputs must be a non-empty mapping of tensors")
        with torch.no_grad():
            hidden = _hidden_states_of(self.encoder(**encoder_inputs))
        return embed_queries(hidden, encoder_inputs.get("attention_mask"), self.pooling,
                             self.normalize, self.query_dtype)

    def search(self, encoder_inputs: Mapping[str, torch.Tensor], indices: Dict[int, Any],
               config: SearchConfig) -> AggregatedSearchResult:
        """Encode -> pool -> normalise -> sharded search -> global top-k; queries never visit the host."""
        return self.aggregator.perform_distributed_search(self.embed(encoder_inputs), indices, config)
