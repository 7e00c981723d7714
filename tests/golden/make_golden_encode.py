"""Golden vectors for the encoder hand-off, produced by importing the REFERENCE's own
``last_token_pool`` (Latest/cuVS-2-gpu/old/generate_embeddings.py:11-21) and applying the
``F.normalize(embeddings, p=2, dim=1)`` of :103, exactly as its ``generate_embeddings_batch`` does.

  encode.npz — per case: hidden [B,T,D] fp32, attention_mask [B,T] int64, pooled = the reference's
               last_token_pool output, normalized = F.normalize(pooled).  Cases: left-padded batch,
               right-padded batch, a batch whose last column is partly masked with one all-zero
               mask row (python's -1 index wrap), all ones, a single sequence, and a zero vector
               (the eps = 1e-12 clamp of F.normalize).

Run: python tests/golden/make_golden_encode.py     (needs /root/reference; not used at test time)
"""
import importlib.util
import os

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Latest/cuVS-2-gpu/old/generate_embeddings.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_generate_embeddings", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)      # defines last_token_pool; main() is guarded
    g = torch.Generator().manual_seed(20261018)
    cases = {}

    def add(name, hidden, mask):
        pooled = ref.last_token_pool(hidden, mask)
        cases[name + "_hidden"] = hidden.numpy()
        cases[name + "_mask"] = mask.numpy()
        cases[name + "_pooled"] = pooled.numpy()
        cases[name + "_normalized"] = F.normalize(pooled, p=2, dim=1).numpy()

    def lengths_mask(lengths, t, left):
        m = torch.zeros((len(lengths), t), dtype=torch.int64)
        for b, n in enumerate(lengths):
            if n:
                if left:
                    m[b, t - n:] = 1
                else:
                    m[b, :n] = 1
        return m

    add("left_padded", torch.randn(5, 9, 48, generator=g), lengths_mask([9, 3, 1, 7, 5], 9, True))
    add("right_padded", torch.randn(6, 11, 40, generator=g), lengths_mask([11, 1, 4, 10, 6, 2], 11, False))
    add("zero_row", torch.randn(4, 7, 24, generator=g), lengths_mask([7, 0, 3, 5], 7, False))
    add("all_ones", torch.randn(3, 5, 384, generator=g), torch.ones(3, 5, dtype=torch.int64))
    add("single", torch.randn(1, 13, 768, generator=g), lengths_mask([8], 13, False))
    h = torch.randn(2, 4, 16, generator=g)
    h[1, 1] = 0.0                                   # the pooled row of sequence 1 is the zero vector
    add("zero_vector", h, lengths_mask([4, 2], 4, False))
    np.savez_compressed(os.path.join(HERE, "encode.npz"), **cases)
    print("wrote encode.npz:", sorted(k for k in cases if k.endswith("_hidden")))


if __name__ == "__main__":
    main()
