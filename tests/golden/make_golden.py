"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the fixtures travel to the GPU box, the reference does not).

  partition.json  — outputs of the REFERENCE's own GPUResourceManager.distribute_workload
                    (imported from /root/reference/Attempt_1) for a sweep of (N, G, strategy).
  knn_l2.npz      — exact L2 neighbours from the reference's CPU baseline path, scikit-learn
                    NearestNeighbors(algorithm='brute') (VectorSearch_QuestionRetrieval.ipynb:L878),
                    on seeded synthetic data.
  knn_ip.npz      — inner-product neighbours by float64 argsort (FAISS IndexFlatIP semantics).
  sample_emb.npz  — the reference's own fixture medical_qa_data/sample_embeddings.pt ([10,384]
                    unit-norm MiniLM rows) and its exact all-pairs top-5.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    sys.path.insert(0, os.path.join(REF, "Attempt_1"))
    from gpu_resource_manager import GPUResourceManager as RefGRM  # the reference implementation

    cases = []
    for n in [1, 2, 3, 7, 10, 100, 300, 301, 302, 1000, 750000, 10_000_000, 100_000_000]:
        for g in [1, 2, 3, 4, 8]:
            m = RefGRM()
            m.available_gpus = list(range(g))
            cases.append({"n": n, "g": g, "strategy": "even",
                          "out": [list(t) for t in m.distribute_workload(n, "even")]})
    m = RefGRM()
    m.available_gpus = [0, 1]
    m.gpu_memory_info = {0: {"available": 8 * 1024 ** 3}, 1: {"available": 16 * 1024 ** 3}}
    cases.append({"n": 300, "g": 2, "strategy": "memory_based", "mem": [8 * 1024 ** 3, 16 * 1024 ** 3],
                  "out": [list(t) for t in m.distribute_workload(300, "memory_based")]})
    with open(os.path.join(HERE, "partition.json"), "w") as f:
        json.dump(cases, f)

    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(20260101)
    db = rng.standard_normal((2000, 64)).astype(np.float32)
    q = rng.standard_normal((40, 64)).astype(np.float32)
    nn = NearestNeighbors(n_neighbors=10, algorithm="brute", metric="euclidean", n_jobs=-1).fit(db)
    d, i = nn.kneighbors(q)
    np.savez_compressed(os.path.join(HERE, "knn_l2.npz"), db=db, q=q, d=(d ** 2).astype(np.float32),
                        i=i.astype(np.int64))
    dbn = db / np.linalg.norm(db, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    sim = qn.astype(np.float64) @ dbn.astype(np.float64).T
    order = np.argsort(-sim, axis=1, kind="stable")[:, :10]
    np.savez_compressed(os.path.join(HERE, "knn_ip.npz"), db=dbn, q=qn,
                        d=np.take_along_axis(sim, order, 1).astype(np.float32), i=order.astype(np.int64))

    emb = torch.load(os.path.join(REF, "Latest/cuVS-2-gpu/medical_qa_data/sample_embeddings.pt"),
                     map_location="cpu").float().numpy()
    sim = emb.astype(np.float64) @ emb.astype(np.float64).T
    order = np.argsort(-sim, axis=1, kind="stable")[:, :5]
    np.savez_compressed(os.path.join(HERE, "sample_emb.npz"), emb=emb,
                        d=np.take_along_axis(sim, order, 1).astype(np.float32), i=order.astype(np.int64))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
