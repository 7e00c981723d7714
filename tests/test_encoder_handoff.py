"""Encoder hand-off: host logic on CPU, kernel parity on the GPU against the
reference's own last_token_pool + F.normalize outputs (tests/golden/encode.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

CASES = ["left_padded", "right_padded", "zero_row", "all_ones", "single", "zero_vector"]


# ------------------------------------------------------------------ host logic (no GPU)
def test_pool_normalize_rejects_bad_arguments(b2):
    with pytest.raises(ValueError, match="3D tensor"):
        b2.pool_normalize(torch.zeros(4, 8))
    with pytest.raises(ValueError, match="CUDA tensors"):
        b2.pool_normalize(torch.zeros(2, 3, 8))          # no CPU path
    with pytest.raises(ValueError, match="unknown pooling"):
        b2.embed_queries(torch.zeros(2, 3, 8), pooling="cls")
    with pytest.raises(ValueError, match="unknown pooling"):
        b2.QueryEncoderHandoff(object(), lambda **kw: None, pooling="max")
    with pytest.raises(TypeError, match="callable"):
        b2.QueryEncoderHandoff(object(), None)


def test_handoff_finds_hidden_states_in_encoder_outputs(b2):
    from types import SimpleNamespace
    eh = b2.encoder_handoff
    h = torch.zeros(1, 2, 4)
    assert eh._hidden_states_of(h) is h
    assert eh._hidden_states_of(SimpleNamespace(last_hidden_state=h)) is h
    assert eh._hidden_states_of({"last_hidden_state": h}) is h
    assert eh._hidden_states_of((h, None)) is h
    with pytest.raises(TypeError):
        eh._hidden_states_of("nope")
    hand = b2.QueryEncoderHandoff(object(), lambda **kw: h)
    with pytest.raises(ValueError, match="non-empty mapping"):
        hand.embed({})


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_last_token_pool_and_normalize_match_reference_golden(b2, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "encode.npz"))
    h = torch.from_numpy(g[case + "_hidden"]).cuda()
    m = torch.from_numpy(g[case + "_mask"]).cuda()
    pooled = b2.last_token_pool(h, m)
    assert pooled.dtype == torch.float32 and pooled.is_cuda
    np.testing.assert_array_equal(pooled.cpu().numpy(), g[case + "_pooled"])        # gather: bit-exact
    out = b2.embed_queries(h, m)
    # fp32 tolerance: the sum of squares is accumulated in a different order than torch's
    np.testing.assert_allclose(out.cpu().numpy(), g[case + "_normalized"], rtol=2e-6, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,out_dtype", [(torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16),
                                             (torch.float16, torch.float32), (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("pooling", ["last_token", "mean"])
def test_pool_normalize_dtypes_against_oracle(b2, dtype, out_dtype, pooling):
    from oracle import encode
    g = torch.Generator().manual_seed(5)
    b, t, d = 7, 33, 200                                     # dim not a multiple of the block
    h = torch.randn(b, t, d, generator=g).to(dtype)
    lengths = torch.tensor([33, 1, 20, 0, 5, 32, 17])
    m = (torch.arange(t)[None, :] < lengths[:, None]).to(torch.int64)
    want = encode.pool_normalize(h.float().numpy(), m.numpy(), pooling)
    got = b2.pool_normalize(h.cuda(), m.cuda(), pooling, True, out_dtype)
    assert got.dtype == out_dtype and tuple(got.shape) == (b, d)
    # one rounding into the output dtype (fp16: 2^-11, bf16: 2^-8 relative) on top of fp32 arithmetic
    tol = {torch.float32: 1e-5, torch.float16: 1e-3, torch.bfloat16: 8e-3}[out_dtype]
    np.testing.assert_allclose(got.float().cpu().numpy(), want, rtol=tol, atol=tol * 0.05)
    # no mask = all ones
    want1 = encode.pool_normalize(h.float().numpy(), None, pooling)
    got1 = b2.pool_normalize(h.cuda(), None, pooling, True, torch.float32)
    np.testing.assert_allclose(got1.cpu().numpy(), want1, rtol=1e-5, atol=1e-7)
    # un-normalised pooling
    want2 = encode.pool_normalize(h.float().numpy(), m.numpy(), pooling, normalize=False)
    got2 = b2.pool_normalize(h.cuda(), m.cuda(), pooling, False, torch.float32)
    np.testing.assert_allclose(got2.cpu().numpy(), want2, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_handoff_searches_without_a_host_hop(b2):
    """encoder -> pool -> normalise -> sharded exact search; the answer equals searching the
    oracle-pooled queries, and the known-answer query (a database row pushed through the
    'encoder') finds itself."""
    from oracle import encode
    g = torch.Generator().manual_seed(11)
    n, d, t = 20000, 128, 6
    db = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    grm = b2.GPUResourceManager(devices=[0])
    a = b2.NativeIndex.flat(db[:12000].to(torch.float16).cuda(), metric="inner_product", id_offset=0)
    c = b2.NativeIndex.flat(db[12000:].to(torch.float16).cuda(), metric="inner_product", id_offset=12000)
    rows = torch.tensor([3, 4567, 12000, 19999])
    hidden = torch.randn(4, t, d, generator=g)
    lengths = torch.tensor([6, 2, 4, 1])
    for bi in range(4):
        hidden[bi, lengths[bi] - 1] = 3.7 * db[rows[bi]]      # the pooled token, un-normalised
    mask = (torch.arange(t)[None, :] < lengths[:, None]).to(torch.int64)
    calls = []

    def encoder(input_ids=None, attention_mask=None):
        calls.append(attention_mask.device.type)
        return {"last_hidden_state": hidden.to(torch.float16).cuda()}

    hand = b2.QueryEncoderHandoff(b2.SearchResultAggregator(grm), encoder, query_dtype=torch.float16)
    inputs = {"input_ids": torch.zeros(4, t, dtype=torch.int64).cuda(), "attention_mask": mask.cuda()}
    q = hand.embed(inputs)
    assert q.is_cuda and q.dtype == torch.float16 and calls == ["cuda"]
    want_q = encode.pool_normalize(hidden.to(torch.float16).float().numpy(), mask.numpy())
    np.testing.assert_allclose(q.float().cpu().numpy(), want_q, rtol=1e-3, atol=1e-4)
    # one GPU holding both shards: search each, merge on the device
    da, ia = a.search(q, 5)
    dc, ic = c.search(q, 5)
    md, mi = b2.merge_topk(torch.stack([da, dc]), torch.stack([ia, ic]), 5, descending=True)
    assert mi[:, 0].cpu().tolist() == rows.tolist()
    assert torch.allclose(md[:, 0].cpu(), torch.ones(4), atol=2e-3)
    res = hand.search(inputs, {0: a}, b2.SearchConfig(k=5))
    assert res.final_indices[0, 0] == 3 and res.final_indices[1, 0] == 4567
    assert res.final_distances.shape == (4, 5)
