"""Golden vectors for the second-generation module (improved_multi_gpu_rag.py), produced by
importing the REFERENCE's own implementation from /root/reference (read-only, dev container only):

  recall.json — RecallEvaluator.calculate_recall_at_k / evaluate_recall_multiple_k outputs on
                seeded id lists, generate_synthetic_ground_truth(5, 1000, 20), the SearchConfig
                defaults and the IndexType values; plus the host merge of ParallelSearchEngine
                (concatenate the shards' (distance, id) lists, argsort, keep top_k — :266-275)
                evaluated on seeded shard results.

Run: python tests/golden/make_golden_recall.py     (needs /root/reference; not used at test time)
"""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Latest/cuVS-2-gpu/improved_multi_gpu_rag.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_improved_multi_gpu_rag", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(7)
    cases = []
    for n_ret, n_rel in [(0, 0), (0, 5), (10, 0), (10, 5), (50, 20), (2000, 100), (7, 100)]:
        retrieved = rng.permutation(3000)[:n_ret]
        relevant = rng.permutation(3000)[:n_rel]
        m = min(n_ret, n_rel // 2)
        if m:
            relevant[:m] = retrieved[:m]          # half of the relevant ids are retrieved
        ks = [1, 5, 10, 50, 100, 500, 1000, 2000]
        cases.append({
            "retrieved": retrieved.tolist(), "relevant": relevant.tolist(), "k_values": ks,
            "single": {str(k): ref.RecallEvaluator.calculate_recall_at_k(retrieved, relevant, k) for k in ks},
            "multi": {str(k): v for k, v in
                      ref.RecallEvaluator.evaluate_recall_multiple_k(retrieved, relevant, ks).items()},
        })
    gt = ref.RecallEvaluator.generate_synthetic_ground_truth(5, 1000, 20)
    cfg = ref.SearchConfig()
    # host merge of ParallelSearchEngine.parallel_search (:266-275) on seeded shard answers
    merges = []
    for shards, kk, top_k in [(2, 8, 5), (3, 6, 6), (8, 10, 10)]:
        d = np.sort(rng.random((shards, kk)).astype(np.float32), axis=1)
        i = rng.permutation(10_000)[: shards * kk].reshape(shards, kk).astype(np.int64)
        all_d, all_i = d.flatten(), i.flatten()
        order = np.argsort(all_d)[:top_k]
        merges.append({"d": d.tolist(), "i": i.tolist(), "top_k": top_k,
                       "out_d": all_d[order].tolist(), "out_i": all_i[order].tolist()})
    out = {
        "recall_cases": cases,
        "synthetic_ground_truth": {str(k): v.tolist() for k, v in gt.items()},
        "search_config": {"top_k": cfg.top_k, "search_batch_size": cfg.search_batch_size,
                          "num_queries": cfg.num_queries, "enable_recall_eval": cfg.enable_recall_eval,
                          "recall_k_values": cfg.recall_k_values},
        "index_types": {t.name: t.value for t in ref.IndexType},
        "merges": merges,
    }
    with open(os.path.join(HERE, "recall.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "recall.json"))


if __name__ == "__main__":
    main()
