"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: -m gpu.

Tolerances (BASELINE.json north_star): BM25 within 1e-5 relative of the rank_bm25 arithmetic;
router gate / fused score within 1e-5 (fp32 inputs) and 1e-3 (bf16-derived inputs); top-k ids
identical modulo ties; integer work (CSR, ids, masks) bit-exact.
"""
import json
import math

from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import bm25_okapi, dense_fusion, philox, router as router_oracle  # noqa: E402


@pytest.fixture(scope="module")
def rq(lib_built):
    import rag_uq_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert rag_uq_b200._lib.lib.ragb_device_check(0) == 0, rag_uq_b200._lib.last_error()
    return rag_uq_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def retrieval_gold(golden_dir):
    with open(golden_dir / "retrieval_golden.json") as fh:
        return json.load(fh)


def assert_ranking_matches(got_ids, got_scores, want_ids, want_scores, all_scores=None, rtol=1e-5, atol=1e-7):
    """"Top-k ids identical modulo ties", made precise: same length, scores equal position by position within the
    tolerance, and every returned id is one the oracle ranks at that position OR a (near-)tie of it - i.e. its
    oracle score is within the tolerance of the oracle score at that position.  ``all_scores`` (the oracle's full
    score vector indexed by id, or a dict id -> score) lets a near-tie ACROSS the cut be recognised; without it
    only the oracle's own top-k can stand in."""
    assert len(got_ids) == len(want_ids), (got_ids, want_ids)
    np.testing.assert_allclose(got_scores, want_scores, rtol=rtol, atol=atol)
    assert len(set(got_ids)) == len(got_ids), "an id was returned twice"
    lookup = dict(zip(want_ids, want_scores))
    for j, (gi, ws) in enumerate(zip(got_ids, want_scores)):
        if gi == want_ids[j]:
            continue
        truth = lookup.get(gi)
        if truth is None and all_scores is not None:
            truth = all_scores[gi] if not isinstance(all_scores, dict) else all_scores.get(gi)
        assert truth is not None, f"rank {j}: id {gi} is not among the oracle's candidates {want_ids}"
        assert abs(truth - ws) <= rtol * abs(ws) + atol, f"rank {j}: id {gi} (oracle score {truth}) is no tie of {want_ids[j]} ({ws})"


# ------------------------------------------------------------------------------------------
# selection
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols,k", [(1, 10, 3), (3, 1000, 10), (2, 70000, 50), (1, 300000, 100), (5, 257, 256),
                                         (64, 4096, 10)])
def test_topk_rows(rq, dev, rows, cols, k):
    g = torch.Generator().manual_seed(rows * 7919 + cols)
    s = torch.randn(rows, cols, generator=g)
    s[:, ::7] = s[:, 3:4]                   # plenty of exact ties
    val, idx = rq.ops.topk_rows(s.to(dev), k)
    want = dense_fusion.topk_desc(s.numpy().astype(np.float64), k)
    for r in range(rows):
        kk = min(k, cols)
        assert idx[r, :kk].tolist() == [w[0] for w in want[r]]
        assert val[r, :kk].tolist() == [np.float32(w[1]) for w in want[r]]
        assert (idx[r, kk:] == -1).all()


def test_topk_rows_ascending_input_worst_case(rq, dev):
    s = torch.arange(50000, dtype=torch.float32).repeat(2, 1)      # every element beats the threshold
    val, idx = rq.ops.topk_rows(s.to(dev), 10)
    assert idx[0].tolist() == list(range(49999, 49989, -1))


def test_topk_merge(rq, dev):
    g = torch.Generator().manual_seed(3)
    b, lists, k_in, k_out = 5, 7, 20, 15
    s = torch.randn(b, lists, k_in, generator=g)
    ids = torch.randperm(b * lists * k_in, generator=g).view(b, lists, k_in).to(torch.int32)
    ids[:, :, -3:] = -1                                             # empty slots
    val, out = rq.ops.topk_merge(s.to(dev), ids.to(dev), k_out)
    for q in range(b):
        parts = [[(int(ids[q, l, j]), float(s[q, l, j])) for j in range(k_in) if ids[q, l, j] >= 0] for l in range(lists)]
        want = dense_fusion.merge_topk(parts, k_out)
        assert out[q].tolist() == [w[0] for w in want]
        assert val[q].tolist() == [np.float32(w[1]) for w in want]


# ------------------------------------------------------------------------------------------
# BM25
# ------------------------------------------------------------------------------------------
def _synthetic_corpus(rq, dev, n, seed_shift=0):
    from rag_uq_b200 import synth
    vocab = synth.vocab_size(n)
    cdf = synth.zipf_cdf(vocab, dev)
    doc_off, doc_tok = synth.doc_tokens(seed_shift, seed_shift + n, cdf)
    return vocab, cdf, doc_off, doc_tok


def test_topk_merge_gathered_reads_the_exchange_buffer_in_place(rq, dev):
    """ragb_topk_merge_strided on the all-gathered exchange layout [ranks, B, 2, 2 * pool] == ragb_topk_merge on the
    transposed contiguous copy the round-1 code made (both sides, empty slots, ties across ranks)."""
    g = torch.Generator().manual_seed(12)
    ranks, n_q, pool = 5, 37, 50
    score = torch.rand(ranks, n_q, 2 * pool, generator=g)
    score[1, :, 10:20] = score[0, :, 10:20]                       # equal scores on two ranks: lower id wins
    ids = torch.randperm(ranks * n_q * 2 * pool, generator=g).view(ranks, n_q, 2 * pool).to(torch.int32)
    ids[2, 3, 40:50] = -1                                         # a short list
    gathered = torch.stack([score.view(torch.int32), ids], dim=2).contiguous().to(dev)      # [G, B, 2, 2 * pool]
    for side in (0, 1):
        got_s, got_i = rq.ops.topk_merge_gathered(gathered, side, pool)
        sl = slice(side * pool, (side + 1) * pool)
        want_s, want_i = rq.ops.topk_merge(score[:, :, sl].permute(1, 0, 2).contiguous().to(dev),
                                           ids[:, :, sl].permute(1, 0, 2).contiguous().to(dev), pool)
        assert torch.equal(got_s, want_s) and torch.equal(got_i, want_i)


def test_reference_checkpoint_loads_through_the_module_alias(rq, dev, golden_dir):
    """A GENUINE reference checkpoint (tests/golden/router_checkpoint.pt: written by the live RouterTrainer.save_checkpoint,
    router.py:499-508, with its pickled RouterConfig, Adam state and loss history) loads through the documented module
    alias, and the B200 router answers what the reference module answered after load_checkpoint - with call-wide statistics
    (stats_initialized is not part of the state dict) and with the running statistics armed."""
    import sys
    import rag_uq_b200.router as router_module
    had = sys.modules.get("rag_uq.router")
    pkg = sys.modules.get("rag_uq")
    sys.modules.setdefault("rag_uq", type(sys)("rag_uq"))
    sys.modules["rag_uq.router"] = router_module
    try:
        ckpt = torch.load(golden_dir / "router_checkpoint.pt", map_location="cpu", weights_only=False)
    finally:
        if had is None:
            sys.modules.pop("rag_uq.router", None)
        else:
            sys.modules["rag_uq.router"] = had
        if pkg is None:
            sys.modules.pop("rag_uq", None)
    assert set(ckpt) == {"model_state_dict", "optimizer_state_dict", "config", "train_losses", "val_losses"}
    assert isinstance(ckpt["config"], rq.RouterConfig) and ckpt["config"].hidden_dim == 32 and ckpt["config"].dropout == 0.2
    router = rq.RetrievalRouter(ckpt["config"])
    router.load_state_dict(ckpt["model_state_dict"])
    router = router.to(dev).eval()
    exp = np.load(golden_dir / "router_checkpoint_expected.npz")
    np.testing.assert_allclose([float(router.bm25_mean), float(router.bm25_std), float(router.dense_mean), float(router.dense_std)],
                               exp["running_stats"], rtol=1e-6)
    b, d = torch.tensor(exp["bm25"], device=dev), torch.tensor(exp["dense"], device=dev)
    with torch.no_grad():
        assert router.stats_initialized is False
        np.testing.assert_allclose(router(b, d).cpu().numpy(), exp["gate_call"], rtol=1e-5, atol=1e-6)
        vals, idx = router.hybrid_rerank(b, d, top_k=10)
        np.testing.assert_allclose(vals.cpu().numpy(), exp["rerank_vals"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(idx.cpu().numpy(), exp["rerank_idx"])
        router.stats_initialized = True
        np.testing.assert_allclose(router(b, d).cpu().numpy(), exp["gate_running"], rtol=1e-5, atol=1e-6)


def test_bm25_known_answers_through_dropin_api(rq, golden_dir):
    with open(golden_dir / "bm25_known_answers.json") as fh:
        known = json.load(fh)
    index = rq.BM25Index()
    index.add_documents([rq.Document(id=f"d{i}", text=t) for i, t in enumerate(known["corpus"])])
    index._ensure_built()
    for word, val in known["idf"].items():
        assert float(index.bm25.idf[index.vocab[word]]) == pytest.approx(val, rel=1e-6, abs=1e-7)
    for query, want in known["queries"].items():
        got = dict(index.search(query, top_k=10))
        expect = {f"d{i}": s for i, s in enumerate(want) if s > 0}
        assert set(got) == set(expect), query
        for key, s in expect.items():
            assert got[key] == pytest.approx(s, rel=1e-5), (query, key)
        ranked = index.search(query, top_k=10)
        assert ranked == sorted(ranked, key=lambda c: (-np.float32(c[1]), int(c[0][1:])))
    assert index.search("zzz") == [] and rq.BM25Index().search("anything") == []


def test_bm25_csr_build_is_bit_exact(rq, dev):
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, 3000)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    ref = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
    assert np.array_equal(shard.term_off.cpu().numpy(), ref.term_off)
    assert np.array_equal(shard.post_doc.cpu().numpy(), ref.post_doc)
    assert np.array_equal(shard.post_tf.cpu().numpy().astype(np.int32), ref.post_tf)
    assert np.array_equal(shard.doc_len.cpu().numpy(), ref.doc_len)
    assert np.array_equal(shard.df.cpu().numpy(), ref.df)
    np.testing.assert_allclose(shard.idf.cpu().numpy(), ref.idf, rtol=2e-7, atol=1e-9)
    norm = 1.5 * (1 - 0.75 + 0.75 * ref.doc_len / ref.avgdl)
    np.testing.assert_allclose(shard.norm.cpu().numpy(), norm, rtol=2e-7)
    # the blocked builder (used for corpora too large to sort at once) gives the same index
    blocks = []
    for lo in range(0, 3000, 700):
        hi = min(3000, lo + 700)
        blocks.append((doc_off[lo:hi + 1] - doc_off[lo], doc_tok[int(doc_off[lo]):int(doc_off[hi])]))
    blocked = rq.build_shard_blocked(iter(blocks), 3000, vocab, dev)
    for name in ("term_off", "post_doc", "post_tf", "doc_len", "df"):
        assert torch.equal(getattr(blocked, name), getattr(shard, name)), name


@pytest.mark.parametrize("table", [True, False])
@pytest.mark.parametrize("n,n_q", [(10_000, 64), (2500, 3), (300, 5)])
def test_bm25_scores_and_topk(rq, dev, n, n_q, table):
    """table=True: frequent terms come from the dense tf table; False: every term streams its postings."""
    from rag_uq_b200 import synth
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    shard = rq.build_shard(doc_off, doc_tok, vocab)
    shard.use_dense_table = table
    shard.finalize()
    assert (shard.dense_terms.numel() > 0) == table
    if table:   # the table is the postings of those terms, byte for byte
        t = int(shard.dense_terms[0])
        a, b = int(shard.term_off[t]), int(shard.term_off[t + 1])
        row = shard.dense_tf[0]
        assert int((row != 0).sum()) == b - a
        assert torch.equal(row[shard.post_doc[a:b].long()].to(torch.int16), shard.post_tf[a:b])
    qb = synth.make_queries(n_q, n, 64, cdf, dev)
    ref = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
    full = shard.scores(qb.q_terms, qb.q_off, qb.max_terms).cpu().numpy()
    terms = qb.q_terms.view(n_q, -1).cpu().numpy()
    for k in (10, 50, 100):
        score, ids = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
        score, ids = score.cpu().numpy(), ids.cpu().numpy()
        for q in range(n_q):
            want_full = ref.get_scores(terms[q])
            if k == 10:
                np.testing.assert_allclose(full[q], want_full, rtol=1e-5, atol=1e-9)   # get_scores parity
            want = bm25_okapi.index_search(want_full, k)
            got = [(int(i), float(s)) for i, s in zip(ids[q], score[q]) if i >= 0]
            assert len(got) == len(want)
            for (gi, gs), (wi, ws) in zip(got, want):
                assert gs == pytest.approx(ws, rel=1e-5)
                assert gi == wi or want_full[gi] == pytest.approx(ws, rel=2e-6), (q, k, gi, wi)
            # the kernel's own order is exactly (fp32 score desc, id asc)
            assert got == sorted(got, key=lambda c: (-c[1], c[0]))


@pytest.mark.parametrize("n,n_q,k", [(300_000, 128, 50), (70_001, 33, 10)])
def test_bm25_impact_bounds_only_prune(rq, dev, n, n_q, k):
    """The fp16 impact bounds of the table terms (tighter table bound + marking pass) must not change a single bit of
    the result: same ids and scores with and without them, and both equal the exhaustive get_scores ranking."""
    from rag_uq_b200 import synth
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    assert shard.dense_imp.numel() > 0 and shard.dense_maximp.shape[0] == shard.dense_terms.shape[0]
    # the bounds really are upper bounds of what the exact path computes
    tf = shard.dense_tf[:, :n].float()
    exact = tf / (tf + shard.norm[None, :])
    assert bool((shard.dense_imp[:, :n].float() >= exact).all())
    assert bool((shard.dense_maximp[:, None] >= exact).all()) and float(shard.dense_maximp.max()) < 1.0
    qb = synth.make_queries(n_q, n, 64, cdf, dev)
    with_s, with_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    keep = shard.dense_imp, shard.dense_maximp
    shard.dense_imp, shard.dense_maximp = keep[0][:0], keep[1][:0]
    without_s, without_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    shard.dense_imp, shard.dense_maximp = keep
    assert torch.equal(with_i, without_i) and torch.equal(with_s, without_s)
    # the impact cap + marker lists (every document above a row's cap is listed; everybody else stays at or below it)
    assert shard.dense_cap.shape[0] == shard.dense_terms.shape[0] and shard.hi_off.shape[0] == shard.dense_terms.shape[0] + 1
    hi_off = shard.hi_off.tolist()
    for r in range(0, shard.dense_terms.shape[0], 7):
        listed = shard.hi_doc[hi_off[r]:hi_off[r + 1]].long()
        assert bool((listed[1:] > listed[:-1]).all())                                     # ascending, distinct
        below = torch.ones(n, dtype=torch.bool, device=dev)
        below[listed] = False
        assert bool((shard.dense_imp[r, :n][below].float() <= shard.dense_cap[r]).all())   # the promise the kernel relies on
        assert bool((exact[r][below] <= shard.dense_cap[r]).all())
    cap = shard.dense_cap, shard.hi_off, shard.hi_doc
    shard.dense_cap = shard.hi_off = shard.hi_doc = None
    nocap_s, nocap_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    shard.dense_cap, shard.hi_off, shard.hi_doc = cap
    assert torch.equal(with_i, nocap_i) and torch.equal(with_s, nocap_s)
    # an aggressive cap (1 % of every row listed) and a degenerate one (a single marker document per row)
    from rag_uq_b200 import sparse as sparse_module
    for tail in (0.01, 1.0 / n):
        old_tail, sparse_module.IMPACT_CAP_TAIL = sparse_module.IMPACT_CAP_TAIL, tail
        shard._build_impact_cap()
        sparse_module.IMPACT_CAP_TAIL = old_tail
        s2, i2 = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
        assert torch.equal(with_i, i2) and torch.equal(with_s, s2)
    shard.dense_cap, shard.hi_off, shard.hi_doc = cap
    full = shard.scores(qb.q_terms, qb.q_off, qb.max_terms)
    want_s, want_i = torch.topk(full, k, dim=1)
    torch.testing.assert_close(with_s, want_s, rtol=0, atol=0)
    same = with_i.long() == want_i
    assert bool((same | (with_s == want_s)).all())      # ids may differ only inside exact score ties


def test_bm25_extreme_document_lengths(rq, dev):
    """rcp.approx (csrc/bm25.cu fast_rcp) outside the synthetic 20-300 token range: one-token passages (norm 0.38 of
    the usual), a 40 000-token passage, term frequencies of 1, 255, 256 (first value that cannot use the byte table),
    3 000 and 60 000 (near the uint16 limit of post_tf) - get_scores within 1e-5 of the float64 rank_bm25 arithmetic
    and the same top-k, with and without the table, with and without baked impacts."""
    g = torch.Generator().manual_seed(99)
    vocab, n = 400, 6_000
    docs = []
    for d in range(n):
        kind = d % 6
        if kind == 0:
            docs.append(torch.randint(0, vocab, (1,), generator=g))                       # one token
        elif kind == 1:
            docs.append(torch.randint(0, vocab, (3,), generator=g))
        else:
            docs.append(torch.randint(0, vocab, (int(torch.randint(20, 300, (1,), generator=g)),), generator=g))
    docs[7] = torch.cat([torch.full((255,), 5), torch.randint(0, vocab, (50,), generator=g)])
    docs[8] = torch.cat([torch.full((256,), 5), torch.randint(0, vocab, (50,), generator=g)])
    docs[9] = torch.cat([torch.full((3000,), 6), torch.full((2000,), 5), torch.randint(0, vocab, (35_000,), generator=g)])
    docs[10] = torch.cat([torch.full((60_000,), 7), torch.randint(0, vocab, (10,), generator=g)])
    docs[11] = torch.full((1,), 7)
    lens = torch.tensor([t.numel() for t in docs])
    doc_off = torch.zeros(n + 1, dtype=torch.int64)
    doc_off[1:] = torch.cumsum(lens, 0)
    doc_tok = torch.cat(docs).to(torch.int32)
    ref = bm25_okapi.OkapiCsr(doc_off.numpy(), doc_tok.numpy(), vocab)
    queries = [[5, 6, 7], [7], [5], [6, 1, 2, 3], [0, 5, 9, 7, 6, 11, 13, 17]]
    q_off = torch.tensor([0] + list(np.cumsum([len(q) for q in queries])), dtype=torch.int32, device=dev)
    q_terms = torch.tensor([t for q in queries for t in q], dtype=torch.int32, device=dev)
    for table in (True, False):
        shard = rq.build_shard(doc_off.to(dev), doc_tok.to(dev), vocab)
        shard.use_dense_table = table
        shard.finalize()
        full = shard.scores(q_terms, q_off, 8).cpu().numpy()
        for qi, q in enumerate(queries):
            want = ref.get_scores(np.asarray(q))
            np.testing.assert_allclose(full[qi], want, rtol=1e-5, atol=1e-9)
        for use_imp in (True, False):
            keep = shard.post_imp
            if not use_imp:
                shard.post_imp = None
            for k in (10, 100):
                score, ids = shard.score_topk(q_terms, q_off, 8, k)
                ts, ti = torch.topk(torch.from_numpy(full).to(dev), k, dim=1)
                ts = torch.where(ts > 0, ts, torch.zeros_like(ts))
                assert torch.equal(score, ts), (table, use_imp, k)
                assert bool(((ids.long() == ti) | (score == 0) | (score == torch.roll(score, 1, 1)) | (score == torch.roll(score, -1, 1))).all())
            shard.post_imp = keep


@pytest.mark.parametrize("n,n_q,k", [(600_000, 256, 50), (300_000, 128, 10), (70_001, 33, 100)])
def test_bm25_baked_impacts_bit_identical(rq, dev, n, n_q, k):
    """Baked impacts (ragb_bm25_build_posting_impacts), read by the window phase instead of tf + norm[doc], are a speed
    device only: post_imp holds the bits the tf + norm path computes, and ids and scores are identical with and
    without them (and equal the exhaustive get_scores ranking)."""
    from rag_uq_b200 import synth
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    assert shard.post_imp is not None and shard.post_imp.shape[0] == shard.nnz
    tf = (shard.post_tf.to(torch.int32) & 0xFFFF).float()
    want_imp = tf / (tf + shard.norm[shard.post_doc.long()])
    torch.testing.assert_close(shard.post_imp, want_imp, rtol=3e-7, atol=0)      # rcp.approx: 1 ulp
    qb = synth.make_queries(n_q, n, 64, cdf, dev)
    with_s, with_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    keep, shard.post_imp = shard.post_imp, None
    without_s, without_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    shard.post_imp = keep
    assert torch.equal(with_i, without_i) and torch.equal(with_s, without_s)
    # also without the impact cap (no marker lists) and without the table (every term is a posting list)
    cap = shard.dense_cap, shard.hi_off, shard.hi_doc
    shard.dense_cap = shard.hi_off = shard.hi_doc = None
    s2, i2 = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    shard.dense_cap, shard.hi_off, shard.hi_doc = cap
    assert torch.equal(with_i, i2) and torch.equal(with_s, s2)
    full = shard.scores(qb.q_terms, qb.q_off, qb.max_terms)
    want_s, want_i = torch.topk(full, k, dim=1)
    want_s = torch.where(want_s > 0, want_s, torch.zeros_like(want_s))
    torch.testing.assert_close(with_s, want_s, rtol=0, atol=0)
    same = (with_i.long() == want_i) | (with_s == 0)
    assert bool((same | (with_s == want_s)).all())


_WINDOW_STRESS = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
import rag_uq_b200 as rq
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
n, vocab, per_doc = 60_000, 96, 12                      # every term occurs in ~12% of the documents
doc_tok = torch.randint(0, vocab, (n * per_doc,), generator=g, dtype=torch.int32).to(dev)
doc_off = (torch.arange(n + 1, dtype=torch.int64) * per_doc).to(dev)
shard = rq.build_shard(doc_off, doc_tok, vocab)
shard.use_dense_table = False                           # all terms stream their posting lists: ub_table = 0, pruned from the start
shard.finalize()
n_q, k = 24, 20
q_terms = torch.randint(0, vocab, (n_q * 30,), generator=g, dtype=torch.int32).to(dev)   # 30 dense lists per query
q_off = (torch.arange(n_q + 1, dtype=torch.int32) * 30).to(dev)
score, ids = shard.score_topk(q_terms, q_off, 30, k)
full = shard.scores(q_terms, q_off, 30)
want_s, want_i = torch.topk(full, k, dim=1)
assert torch.equal(score, want_s), (score - want_s).abs().max()
assert bool(((ids.long() == want_i) | (score == want_s)).all())
print("window stress ok")
"""


@pytest.mark.parametrize("target", ["100000", "256", "0"])
def test_bm25_window_mode_overflow_and_fallback(rq, dev, target):
    """Window mode under stress: 30 dense posting lists per query overflow the 512-slot hash table at any window size
    (halving down to one super-range, then back to the dense accumulator); an absurd posting target starts every
    window at 32k documents.  The results must equal the exhaustive ranking bit for bit in every setting
    (RAGB_BM25_WINDOW is read once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, RAGB_BM25_WINDOW=target)
    root = str(Path(__file__).resolve().parent.parent)
    out = subprocess.run([sys.executable, "-c", _WINDOW_STRESS, root], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "window stress ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("seed", list(range(10)))
def test_bm25_random_small_corpora_vs_oracle(rq, dev, seed):
    """Differential test on odd shapes: random tiny / ragged corpora (empty documents, vocabularies from 3 to 5000
    terms, with and without the byte table), random k, queries with duplicates and out-of-vocabulary ids - top-k
    scores must equal the float64 oracle within 1e-5 and the exhaustive GPU ranking bit for bit."""
    g = torch.Generator().manual_seed(1000 + seed)
    n = int(torch.randint(1, 3000, (1,), generator=g))
    vocab = int(torch.randint(3, 5000, (1,), generator=g))
    lens = torch.randint(0, 40, (n,), generator=g)
    if seed % 3 == 0:
        lens[torch.randint(0, n, (max(1, n // 7),), generator=g)] = 0          # empty documents
    doc_off = torch.zeros(n + 1, dtype=torch.int64)
    doc_off[1:] = torch.cumsum(lens, 0)
    skew = torch.rand(int(doc_off[-1]), generator=g) ** (2 + seed % 4)          # a few very frequent terms
    doc_tok = (skew * vocab).to(torch.int32).clamp_(max=vocab - 1)
    shard = rq.build_shard(doc_off.to(dev), doc_tok.to(dev), vocab)
    shard.use_dense_table = seed % 2 == 0
    shard.finalize()
    n_q = int(torch.randint(1, 40, (1,), generator=g))
    q_len = torch.randint(0, 12, (n_q,), generator=g)
    q_off = torch.zeros(n_q + 1, dtype=torch.int32)
    q_off[1:] = torch.cumsum(q_len, 0).to(torch.int32)
    q_terms = torch.randint(-2, vocab + 3, (max(1, int(q_off[-1])),), generator=g, dtype=torch.int32)   # some OOV ids
    k = int(torch.randint(1, 120, (1,), generator=g))
    max_terms = max(1, int(q_len.max()))
    score, ids = shard.score_topk(q_terms.to(dev), q_off.to(dev), max_terms, k)
    full = shard.scores(q_terms.to(dev), q_off.to(dev), max_terms)
    kk = min(k, n)
    want_s, want_i = torch.topk(full, kk, dim=1)
    want_s = torch.where(want_s > 0, want_s, torch.zeros_like(want_s))          # only positive scores are returned
    assert torch.equal(score[:, :kk], want_s)
    assert bool((score[:, kk:] == 0).all()) and bool((ids[:, kk:] == -1).all())
    assert bool(((ids[:, :kk].long() == want_i) | (want_s == 0) | (score[:, :kk] == want_s)).all())
    assert bool((ids[:, :kk][want_s == 0] == -1).all())
    ref = bm25_okapi.OkapiCsr(doc_off.numpy(), doc_tok.numpy(), vocab)
    for q in range(n_q):
        terms = q_terms[int(q_off[q]):int(q_off[q + 1])].tolist()
        np.testing.assert_allclose(full[q].cpu().numpy(), ref.get_scores(terms), rtol=1e-5, atol=1e-9)


def test_bm25_edge_queries(rq, dev):
    """empty query, all-OOV query, duplicated terms, a query longer than the corpus is wide."""
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, 1000)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    ref = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
    queries = [[], [vocab + 3, -1], [5, 5, 5, 17], list(range(40, 100))]
    flat = torch.tensor([t for q in queries for t in q] or [0], dtype=torch.int32, device=dev)
    off = torch.tensor(np.concatenate([[0], np.cumsum([len(q) for q in queries])]), dtype=torch.int32, device=dev)
    score, ids = shard.score_topk(flat, off, 60, 20)
    full = shard.scores(flat, off, 60).cpu().numpy()
    assert (ids[0] == -1).all() and (ids[1] == -1).all() and (score[:2] == 0).all()
    for q in (2, 3):
        want_full = ref.get_scores(queries[q])
        np.testing.assert_allclose(full[q], want_full, rtol=1e-5, atol=1e-9)
        want = bm25_okapi.index_search(want_full, 20)
        got_ids = [int(i) for i in ids[q] if i >= 0]
        got_scores = [float(s) for s, i in zip(score[q], ids[q]) if i >= 0]
        assert_ranking_matches(got_ids, got_scores, [w[0] for w in want], [w[1] for w in want], want_full, 1e-5)


def test_bm25_shards_with_global_statistics_equal_unsharded(rq, dev):
    """Two document shards scored with global df / N / avgdl and merged == one index."""
    from rag_uq_b200 import synth
    n, k = 6000, 50
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    whole = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    qb = synth.make_queries(16, n, 64, cdf, dev)
    ws, wi = whole.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    parts = []
    bounds = [rq.shard_rows(n, 2, r) for r in range(2)]
    shards = []
    for lo, hi in bounds:
        off = doc_off[lo:hi + 1] - doc_off[lo]
        tok = doc_tok[int(doc_off[lo]):int(doc_off[hi])]
        shards.append(rq.build_shard(off, tok, vocab, id_base=lo))
    df = shards[0].df + shards[1].df
    total_len = int(shards[0].doc_len.sum() + shards[1].doc_len.sum())
    for sh in shards:
        sh.finalize(df, n, total_len)
        parts.append(sh.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k))
    s = torch.stack([p[0] for p in parts], dim=1)
    i = torch.stack([p[1] for p in parts], dim=1)
    ms, mi = rq.ops.topk_merge(s, i, k)
    assert torch.equal(mi, wi) and torch.equal(ms, ws)            # bit-identical, not just close
    # the same with the pruning bounds exchanged between the shards (what HybridEngine does at world > 1): every
    # shard starts from the MAXIMUM of the shards' proven bounds - still bit-identical to the unsharded index
    seeds = torch.stack([sh.seed(qb.q_terms, qb.q_off, qb.max_terms, k) for sh in shards]).amax(dim=0)
    parts = [sh.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k, seeds) for sh in shards]
    ms2, mi2 = rq.ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1), k)
    assert torch.equal(mi2, wi) and torch.equal(ms2, ws)
    # ... and with every shard seeding only ITS slice of the batch (q_off sliced, offsets absolute), the others' entries 0:
    # the MAX then hands each query the bound of the one shard that seeded it (HybridEngine.local_pools, world > 1)
    n_q = qb.q_off.shape[0] - 1
    per = -(-n_q // len(shards))
    sliced = torch.zeros(n_q, dtype=torch.float32, device=dev)
    for r, sh in enumerate(shards):
        q0, q1 = min(n_q, r * per), min(n_q, (r + 1) * per)
        part = sh.seed(qb.q_terms, qb.q_off[q0:q1 + 1], qb.max_terms, k)
        full = sh.seed(qb.q_terms, qb.q_off, qb.max_terms, k)
        assert torch.equal(part, full[q0:q1])                      # a sliced call seeds exactly those queries
        sliced[q0:q1] = part
    assert bool((sliced <= seeds).all()) and bool((sliced > 0).any())
    parts = [sh.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k, sliced) for sh in shards]
    ms3, mi3 = rq.ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1), k)
    assert torch.equal(mi3, wi) and torch.equal(ms3, ws)


def test_bm25_staged_search_and_two_stream_overlap_equal_the_serial_path(rq, dev):
    """ragb_bm25_score_part / _finish (stripes scored in several launches, optionally with padded shared memory) equal
    ragb_bm25_score_topk bit for bit; HybridEngine.local_pools(overlap=True) - BM25 blocks co-resident with the 4-stage
    tcgen05 kernel on a second, higher-priority stream - returns exactly what the serial path returns."""
    from rag_uq_b200 import synth
    n, n_q, k = 1_000_000, 300, 50
    engine, cdf = synth.build_synthetic_engine(n, 768, dev)
    sh = engine.sparse
    qb = synth.make_queries(n_q, n, 768, cdf, dev)
    want_s, want_i = sh.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    stripes = rq.ops.bm25_stripe_count(n_q, n)
    assert stripes >= 3
    for cuts, pad in (([0, 1, stripes], 0), ([0, stripes // 2, stripes - 1, stripes], 96 * 1024), ([0, stripes], 0)):
        ws = rq.ops.bm25_workspace(n_q, n, k, dev)
        for a, b in zip(cuts[:-1], cuts[1:]):
            sh.score_part(qb.q_terms, qb.q_off, qb.max_terms, k, ws, a, b, pad if a == 0 else 0)
        s, i = rq.ops.bm25_score_finish(n_q, n, k, ws)
        assert torch.equal(i, want_i) and torch.equal(s, want_s)
    with pytest.raises(ValueError):
        sh.score_part(qb.q_terms, qb.q_off, qb.max_terms, k, rq.ops.bm25_workspace(n_q, n, k, dev), 2, stripes + 1)
    serial = engine.local_pools(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, k, overlap=False)
    for _ in range(3):
        both = engine.local_pools(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, k, overlap=True)
        torch.cuda.synchronize()
        for a, b in zip(both, serial):
            assert torch.equal(a, b)
    # variant 4 (CTA pairs, 4-stage ring) == variant 3
    s3, i3 = rq.ops.dense_mma_topk(engine.passages, qb.q_emb, k, 0, 3)
    s4, i4 = rq.ops.dense_mma_topk(engine.passages, qb.q_emb, k, 0, 4)
    assert torch.equal(s3, s4) and torch.equal(i3, i4)


@pytest.mark.parametrize("n,n_q,k", [(120_000, 48, 50), (9_000, 20, 10)])
def test_bm25_external_seed_only_prunes(rq, dev, n, n_q, k):
    """ragb_bm25_seed + seed_thr: any PROVEN lower bound of the k-th best score gives the same result - the kernel's own
    seed, no seed at all (zeros), and the tightest possible one (the true k-th best score itself)."""
    from rag_uq_b200 import synth
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    qb = synth.make_queries(n_q, n, 64, cdf, dev)
    base_s, base_i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    own = shard.seed(qb.q_terms, qb.q_off, qb.max_terms, k)
    kth = torch.where(base_i[:, k - 1] >= 0, base_s[:, k - 1], torch.zeros_like(base_s[:, k - 1]))
    assert bool((own <= kth + 1e-6).all()), "a seed exceeds the k-th best score it is supposed to bound from below"
    for seed in (own, torch.zeros_like(own), kth.contiguous()):
        s, i = shard.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k, seed)
        assert torch.equal(i, base_i) and torch.equal(s, base_s)


# ------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------
def _dense_case(rq, dev, n, b, dim=768):
    from rag_uq_b200 import synth
    cdf = synth.zipf_cdf(synth.vocab_size(n), dev)
    passages = synth.passage_embeddings(0, n, dim, dev)
    qb = synth.make_queries(b, n, dim, cdf, dev)
    want = dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy())
    return passages, qb.q_emb, want


def _check_dense(score, ids, want_scores, k, id_base=0):
    want = dense_fusion.topk_desc(want_scores, k)
    score, ids = score.cpu().numpy(), ids.cpu().numpy()
    for q in range(want_scores.shape[0]):
        for j, (wi, ws) in enumerate(want[q]):
            gi, gs = int(ids[q, j]) - id_base, float(score[q, j])
            assert gs == pytest.approx(ws, abs=2e-6), (q, j)
            assert gi == wi or want_scores[q, gi] == pytest.approx(ws, abs=2e-6), (q, j, gi, wi)


@pytest.mark.parametrize("n,b,k,dim", [(10_000, 1, 10, 768), (5000, 3, 50, 768), (4097, 8, 100, 768), (700, 2, 10, 384),
                                       (900, 1, 256, 64)])
def test_dense_gemv_topk(rq, dev, n, b, k, dim):
    passages, q, want = _dense_case(rq, dev, n, b, dim)
    score, ids = rq.ops.dense_gemv_topk(passages, q, k, 1000)
    _check_dense(score, ids, want, k, id_base=1000)
    full = rq.ops.dense_scores(passages, q).cpu().numpy()
    np.testing.assert_allclose(full, want, atol=2e-6)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n,b,k,dim", [(10_000, 64, 10, 768), (4096, 128, 50, 768), (33_333, 200, 10, 768),
                                       (1000, 9, 100, 768), (2048, 300, 50, 128), (127, 130, 10, 768)])
def test_dense_mma_topk(rq, dev, variant, n, b, k, dim):
    passages, q, want = _dense_case(rq, dev, n, b, dim)
    score, ids = rq.ops.dense_mma_topk(passages, q, k, 7, variant)
    torch.cuda.synchronize()
    _check_dense(score, ids, want, min(k, n), id_base=7)


@pytest.mark.parametrize("n,b,k,variant", [(150_000, 200, 50, 3), (140_000, 130, 10, 2), (131_072, 300, 100, 3)])
def test_dense_mma_seeded_two_phase(rq, dev, n, b, k, variant):
    """Shards of >= 512 passage tiles are searched in two phases (sampled prefix -> proven bound -> seeded rest):
    the result equals the float64 oracle; the explicit two-call form (ragb_dense_mma_sample / _seeded) equals the
    one-call form bit for bit - with its own bounds, with no bounds (-inf) and with the tightest valid bounds (the true
    k-th best scores, as another shard might have proven them)."""
    passages, q, want = _dense_case(rq, dev, n, b)
    score, ids = rq.ops.dense_mma_topk(passages, q, k, 11, variant)
    _check_dense(score, ids, want, k, id_base=11)
    thr, ws = rq.ops.dense_mma_sample(passages, q, k, 11, variant)
    assert bool(torch.isfinite(thr).all()) and bool((thr <= score[:, k - 1]).all())     # proven LOWER bounds
    s2, i2 = rq.ops.dense_mma_seeded(passages, q, k, 11, variant, thr, ws)
    assert torch.equal(s2, score) and torch.equal(i2, ids)
    for bound in (torch.full_like(thr, float("-inf")), score[:, k - 1].contiguous()):
        thr3, ws3 = rq.ops.dense_mma_sample(passages, q, k, 11, variant)
        s3, i3 = rq.ops.dense_mma_seeded(passages, q, k, 11, variant, torch.maximum(thr3, bound), ws3)
        assert torch.equal(s3, score) and torch.equal(i3, ids)
    # small shards have no sampled prefix: the bound is -inf and the seeded phase covers everything
    p_small, q_small, want_small = _dense_case(rq, dev, 5000, 130)
    thr, ws = rq.ops.dense_mma_sample(p_small, q_small, k, 0, variant)
    assert bool(torch.isinf(thr).all()) and bool((thr < 0).all())
    s4, i4 = rq.ops.dense_mma_seeded(p_small, q_small, k, 0, variant, thr, ws)
    _check_dense(s4, i4, want_small, min(k, 5000))


def test_dense_mma_topk_min_reports_a_lower_bound_of_the_smallest_score(rq, dev):
    passages, q, want = _dense_case(rq, dev, 140_000, 200)
    s0, i0 = rq.ops.dense_mma_topk(passages, q, 10, 0, 3)
    s1, i1, lowest = rq.ops.dense_mma_topk_min(passages, q, 10, 0, 3)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    true_min = want.min(axis=1)
    got = lowest.cpu().numpy()
    assert (got <= true_min + 2e-6).all() and (got >= np.minimum(true_min, 0.0) - 2e-6).all()
    assert np.abs(got - true_min).max() <= 2e-6          # 140000 = 546.9 tiles: the padded rows (score 0) do not undercut


def test_hnsw_recall_against_exact_search(rq, dev):
    """The reference's dense path is ChromaDB's approximate HNSW (streaming_index.py:355-359); ours is exact.
    Report (and sanity-check) the recall@10 of an HNSW with ChromaDB's default parameters against the kernel."""
    from oracle import hnsw
    from rag_uq_b200 import synth
    n, n_q, k = 3000, 48, 10
    passages = synth.passage_embeddings(0, n, 768, dev)
    qb = synth.make_queries(n_q, n, 768, synth.zipf_cdf(synth.vocab_size(n), dev), dev)
    score, ids = rq.ops.dense_gemv_topk(passages, qb.q_emb[:8].contiguous(), k, 0)
    score2, ids2 = rq.ops.dense_mma_topk(passages, qb.q_emb, k, 0, 3)
    assert torch.equal(ids, ids2[:8])
    rep = hnsw.recall_report(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy(), ids2.cpu().numpy(), k)
    print("HNSW recall vs exact:", rep)
    # the source passage of each query (cosine ~0.9, everything else ~0.0 +- 0.04) is always found ...
    index = hnsw.HnswCosine(768)
    index.add(passages.float().cpu().numpy())
    top1 = [int(index.search(q, 1, 100)[0][0]) for q in qb.q_emb.float().cpu().numpy()]
    assert np.mean(np.asarray(top1) == ids2[:, 0].cpu().numpy()) > 0.95
    # ... the other nine are near-random directions in 768-d, where a graph walk is genuinely approximate
    assert 0.3 < rep["recall@10_search_ef_10"] <= rep["recall@10_search_ef_100"] <= 1.0


def test_dense_gemv_equals_mma_bitwise_ids(rq, dev):
    passages, q, want = _dense_case(rq, dev, 20_000, 8)
    gs, gi = rq.ops.dense_gemv_topk(passages, q, 50, 0)
    ms, mi = rq.ops.dense_mma_topk(passages, q, 50, 0, 0)
    agree = (gi == mi).float().mean().item()
    assert agree > 0.99                                              # accumulation order differs, ids should not
    torch.testing.assert_close(gs, ms, atol=2e-6, rtol=0)


# ------------------------------------------------------------------------------------------
# fusion + end to end (config C1: 10k passages x 768, 64 queries, top-10, pool 50)
# ------------------------------------------------------------------------------------------
def test_hybrid_fuse_topk_vs_oracle(rq, dev):
    g = torch.Generator().manual_seed(17)
    b, pool, k = 9, 50, 10
    bs = torch.rand(b, pool, generator=g) * 20
    ds = torch.rand(b, pool, generator=g) * 2 - 0.8
    bi = torch.stack([torch.randperm(120, generator=g)[:pool] for _ in range(b)]).to(torch.int32)
    di = torch.stack([torch.randperm(120, generator=g)[:pool] for _ in range(b)]).to(torch.int32)
    bs, _ = torch.sort(bs, dim=1, descending=True)
    ds, _ = torch.sort(ds, dim=1, descending=True)
    bi[2, 30:] = -1; bs[2, 30:] = 0                 # short BM25 pool
    bi[3, :] = -1; bs[3, :] = 0                     # no BM25 hit at all
    ds[4, :] = -torch.rand(pool, generator=g).sort().values   # all-negative dense pool
    bi[5, :] = -1; di[5, :] = -1                    # nothing at all
    ids, ob, od, oh = (t.cpu() for t in rq.ops.hybrid_fuse_topk(bs.to(dev), bi.to(dev), ds.to(dev), di.to(dev), k))
    for q in range(b):
        bm = [(int(i), float(s)) for i, s in zip(bi[q], bs[q]) if i >= 0]
        de = [(int(i), float(s)) for i, s in zip(di[q], ds[q]) if i >= 0]
        want = dense_fusion.hybrid_search(bm, de, k)
        got = [int(i) for i in ids[q] if i >= 0]
        assert got == [w[0] for w in want], q
        np.testing.assert_allclose(oh[q, :len(want)], [w[3] for w in want], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ob[q, :len(want)], [w[1] for w in want], rtol=0, atol=0)
        np.testing.assert_allclose(od[q, :len(want)], [w[2] for w in want], rtol=0, atol=0)
        assert (ids[q, len(want):] == -1).all() and (oh[q, len(want):] == 0).all()


@pytest.mark.parametrize("n_q,k,pool", [(64, 10, 50), (4, 10, 50), (40, 100, 100)])
def test_hybrid_engine_end_to_end_c1(rq, dev, n_q, k, pool):
    """Config C1 end to end against the oracle; (40, 100, 100) is the top-100 / pool-100 shape of config C5."""
    from rag_uq_b200 import synth
    n, dim = 10_000, 768
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, dim, dev)
    shard = rq.build_shard(doc_off, doc_tok, vocab).finalize()
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    engine = rq.HybridEngine(shard, passages)
    ids, sb, sd, sh = (t.cpu().numpy() for t in engine.hybrid_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, k, pool))
    okapi = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
    dense = dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy())
    terms = qb.q_terms.view(n_q, -1).cpu().numpy()
    exact = 0
    for q in range(n_q):
        bm = bm25_okapi.index_search(okapi.get_scores(terms[q]), pool)
        de = dense_fusion.topk_desc(dense[q:q + 1], pool)[0]
        want = dense_fusion.hybrid_search(bm, de, k)
        got = [int(i) for i in ids[q] if i >= 0]
        exact += got == [w[0] for w in want]
        # every deviation must be a near-tie in the ORACLE's own hybrid scores (pool-boundary and rank ties only)
        everything = {w[0]: w[3] for w in dense_fusion.hybrid_search(
            bm25_okapi.index_search(okapi.get_scores(terms[q]), pool + 3), dense_fusion.topk_desc(dense[q:q + 1], pool + 3)[0], 4 * pool)}
        assert_ranking_matches(got, sh[q, :len(got)].tolist(), [w[0] for w in want], [w[3] for w in want], everything, 2e-5, 1e-6)
    assert exact >= n_q - max(1, n_q // 20)   # and they are rare
    # router-in-the-loop + confidence (run_evaluation.py:165-196) on the same batch
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    with torch.no_grad():
        res = engine.retrieve_and_rerank(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, pool, mc_samples=10, seed=5)
    state = {key: v.detach().cpu() for key, v in router.state_dict().items()}
    # the reference loop (run_evaluation.py:165-184) calls the router once per query with [1, k] tensors: statistics of
    # THAT query only (stats_initialized is False right after load_state_dict), never of its batch mates
    per_query = [router_oracle.hybrid_rerank(torch.tensor(sb[q:q + 1]), torch.tensor(sd[q:q + 1]), state, False, k) for q in range(n_q)]
    ov, oi = torch.cat([p[0] for p in per_query]), torch.cat([p[1] for p in per_query])
    torch.testing.assert_close(res["fused"].cpu(), ov, rtol=1e-5, atol=1e-5)
    want_ids = np.take_along_axis(ids, oi.numpy(), axis=1)
    got_ids, got_vals = res["ids"].cpu().numpy(), res["fused"].cpu().numpy()
    for q in range(n_q):      # padding entries (id -1) all fuse to the same value: their mutual order is a tie
        real = want_ids[q] >= 0
        every = {int(i): float(v) for i, v in zip(want_ids[q][real], ov[q].numpy()[real])}
        assert_ranking_matches([int(i) for i in got_ids[q] if i >= 0], [float(v) for v, i in zip(got_vals[q], got_ids[q]) if i >= 0],
                               [int(i) for i in want_ids[q][real]], [float(v) for v in ov[q].numpy()[real]], every, 1e-5, 1e-5)
    if n_q > 1:   # ... and it differs from normalising over the whole [B, k] call, which hybrid_rerank does by default
        bv, bi_ = router_oracle.hybrid_rerank(torch.tensor(sb), torch.tensor(sd), state, False, k)
        dv, di_ = router.hybrid_rerank(torch.tensor(sb, device=dev), torch.tensor(sd, device=dev), top_k=k)
        torch.testing.assert_close(dv.cpu(), bv, rtol=1e-5, atol=1e-5)
        assert not torch.allclose(bv, ov)
    for q in range(n_q):
        valid = [float(v) for v, i in zip(res["fused"][q], res["ids"][q]) if i >= 0]
        assert float(res["retrieval_uncertainty"][q]) == pytest.approx(dense_fusion.retrieval_uncertainty(valid, 1.0), rel=1e-4, abs=1e-5)
    assert res["router_confidence"].shape == (n_q,) and bool(((res["router_confidence"] >= 0) & (res["router_confidence"] <= 1)).all())
    # the source passage of every query must be its best dense hit (sanity of the synthetic design)
    ds, di = engine.dense_topk(qb.q_emb, 1)
    assert (di[:, 0].cpu() == qb.source_rows.cpu().to(torch.int32)).float().mean() > 0.95


def test_full_fusion_mode_c1(rq, dev):
    """Config C1 in full-fusion mode: gate on every (query, passage) pair, oracle = hybrid_rerank on [B, N]."""
    from rag_uq_b200 import synth
    n, dim, k, n_q = 10_000, 768, 10, 64
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, dim, dev)
    engine = rq.HybridEngine(rq.build_shard(doc_off, doc_tok, vocab).finalize(), passages)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    with pytest.raises(ValueError):
        engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k)
    router.bm25_mean.fill_(6.0); router.bm25_std.fill_(5.0); router.dense_mean.fill_(0.0); router.dense_std.fill_(0.05)
    router.stats_initialized = True
    with torch.no_grad():
        score, ids = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, query_chunk=24)
    okapi = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
    terms = qb.q_terms.view(n_q, -1).cpu().numpy()
    bm = torch.tensor(np.stack([okapi.get_scores(terms[q]) for q in range(n_q)]), dtype=torch.float32)
    de = torch.tensor(dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy()),
                      dtype=torch.float32)
    state = {key: v.detach().cpu() for key, v in router.state_dict().items()}
    want_s, want_i = router_oracle.hybrid_rerank(bm, de, state, True, k)
    torch.testing.assert_close(score.cpu(), want_s, rtol=2e-5, atol=2e-5)
    assert (ids.cpu().long() == want_i).float().mean() > 0.99


@pytest.mark.parametrize("n,n_q,k,hidden,scale", [(10_000, 64, 10, 64, 1.0), (33_331, 200, 50, 32, 4.0),
                                                  (200_000, 384, 10, 64, 1.0), (4099, 130, 100, 128, 10.0)])
def test_full_fusion_fused_epilogue(rq, dev, n, n_q, k, hidden, scale):
    """The tcgen05 epilogue with gate + fusion inside equals the un-fused full-fusion path (itself pinned to the
    oracle in test_full_fusion_mode_c1) and the torch-CPU oracle; the gate-bound table prunes without changing results."""
    from rag_uq_b200 import synth
    dim = 768
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, dim, dev)
    engine = rq.HybridEngine(rq.build_shard(doc_off, doc_tok, vocab).finalize(), passages, id_base=1000)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    torch.manual_seed(11)
    router = rq.RetrievalRouter(rq.RouterConfig(hidden_dim=hidden)).to(dev).eval()
    with torch.no_grad():
        for p in router.parameters():
            p.mul_(scale)                      # scale > 1: a gate that swings between 0 and 1 like a trained one
    router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.1); router.dense_std.fill_(0.2)
    router.stats_initialized = True
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.no_grad():
        fs, fi = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=True,
                                         counters=counters)
        us, ui = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=False,
                                         query_chunk=64)
    # the two paths see dense scores from different fp32 accumulation orders (tensor core vs CUDA cores, ~1e-7);
    # a steep gate (scale > 1) amplifies that by |dense - bm25| * gate slope
    tol = 1e-5 * max(1.0, scale * scale)
    torch.testing.assert_close(fs, us, rtol=tol, atol=tol)
    same = (fi == ui)
    assert same.float().mean() > 0.99
    # where the ids differ the scores must be within rounding of each other (a near-tie swapped)
    assert bool(((fs - us).abs()[~same] <= tol * us.abs()[~same] + 1e-6).all())
    evals, admits = (int(v) for v in counters.tolist())
    assert 0 < admits <= evals <= n_q * n
    if n >= 100_000:                           # the bound prunes most gate evaluations once the lists have warmed up
        assert evals < 0.2 * n_q * n           # (a small corpus never leaves warm-up: ~150 passages per list)
    if n <= 40_000:
        okapi = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
        terms = qb.q_terms.view(n_q, -1).cpu().numpy()
        bm = torch.tensor(np.stack([okapi.get_scores(terms[q]) for q in range(n_q)]), dtype=torch.float32)
        de = torch.tensor(dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy()),
                          dtype=torch.float32)
        state = {key: v.detach().cpu() for key, v in router.state_dict().items()}
        want_s, want_i = router_oracle.hybrid_rerank(bm, de, state, True, k)
        torch.testing.assert_close(fs.cpu(), want_s, rtol=3 * tol, atol=3 * tol)
        assert (fi.cpu().long() - 1000 == want_i).float().mean() > 0.98


def test_score_docs_kernels(rq, dev):
    """ragb_bm25_score_docs == the streaming kernel's get_scores at the chosen documents, BIT FOR BIT (same arithmetic,
    same summation order; duplicated query terms, OOV ids, ids outside the shard, -1 pads); ragb_dense_score_docs == the
    float64 inner product to fp32 rounding."""
    from rag_uq_b200 import synth
    n, dim, n_q, c = 20_000, 768, 24, 37
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    shard = rq.build_shard(doc_off, doc_tok, vocab, id_base=500).finalize()
    passages = synth.passage_embeddings(0, n, dim, dev)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    terms = qb.q_terms.clone().view(n_q, -1)
    terms[3, 4:] = terms[3, :4]                       # duplicated terms count per occurrence
    terms[5, 0] = vocab + 9                           # out of vocabulary
    q_terms = terms.reshape(-1).contiguous()
    full = shard.scores(q_terms, qb.q_off, qb.max_terms)                      # [n_q, n] from bm25_kernel<DENSE_OUT>
    g = torch.Generator(device="cpu").manual_seed(3)
    cand = torch.randint(0, n, (n_q, c), generator=g).to(dev).to(torch.int32) + 500
    cand[:, 0] = qb.source_rows.to(torch.int32) + 500                         # a document that matches every list term
    cand[2, 5], cand[7, 1], cand[9, 2] = -1, 499, 500 + n                     # pad, below the shard, above the shard
    got = shard.score_docs(q_terms, qb.q_off, qb.max_terms, cand.contiguous())
    local = (cand - 500).long()
    valid = (local >= 0) & (local < n)
    want = torch.where(valid, torch.gather(full, 1, local.clamp(0, n - 1)), torch.zeros((), device=dev))
    assert torch.equal(got, want)
    # and against the top-k kernel's own scores
    ts, ti = shard.score_topk(q_terms, qb.q_off, qb.max_terms, 50)
    again = shard.score_docs(q_terms, qb.q_off, qb.max_terms, ti)
    assert torch.equal(torch.where(ti >= 0, again, torch.zeros_like(again)), torch.where(ti >= 0, ts, torch.zeros_like(ts)))
    d = rq.ops.dense_score_docs(passages, qb.q_emb, 500, cand.contiguous())
    want_d = dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy())
    want_d = np.where(valid.cpu().numpy(), np.take_along_axis(want_d, local.clamp(0, n - 1).cpu().numpy(), axis=1), 0.0)
    np.testing.assert_allclose(d.cpu().numpy(), want_d, rtol=0, atol=2e-6)


@pytest.mark.parametrize("n,n_q,k,hidden,scale,dense_favoured", [(10_000, 64, 10, 64, 1.0, False), (33_331, 200, 50, 32, 4.0, False),
                                                                 (20_000, 150, 10, 16, 8.0, True), (150_000, 256, 10, 64, 1.0, False)])
def test_full_fusion_threshold_algorithm(rq, dev, n, n_q, k, hidden, scale, dense_favoured):
    """Full-fusion as a threshold-algorithm search (no [B, N] matrix): equals RetrievalRouter.hybrid_rerank over ALL
    passages (torch-CPU oracle where the corpus is small enough, the exhaustive epilogue otherwise) - with the default
    depth (stopping rule proven for nearly every query) and with a depth so small that most queries take the fallback."""
    from rag_uq_b200 import synth
    dim = 768
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, dim, dev)
    engine = rq.HybridEngine(rq.build_shard(doc_off, doc_tok, vocab, id_base=1000).finalize(), passages, id_base=1000)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    torch.manual_seed(11)
    router = rq.RetrievalRouter(rq.RouterConfig(hidden_dim=hidden)).to(dev).eval()
    with torch.no_grad():
        for p in router.parameters():
            p.mul_(scale)
        if dense_favoured:
            router.scorer[3].bias.fill_(6.0)         # gate ~ 1 almost everywhere: fused ~ dense, the BM25 order says little
    router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.1); router.dense_std.fill_(0.2)
    router.stats_initialized = True
    tol = 1e-5 * max(1.0, scale * scale)
    with torch.no_grad():
        info, info_small = {}, {}
        ts, ti = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, method="threshold", info=info)
        ss, si = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, method="threshold",
                                         depth=k, info=info_small)
        es, ei = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, method="exhaustive")
    assert info["method"] == "threshold" and info_small["fallback_queries"] >= info["fallback_queries"]
    if scale == 1.0:
        assert info["fallback_queries"] <= n_q // 4      # a smooth gate: the stopping rule holds for most queries at depth 100
    for got_s, got_i in ((ts, ti), (ss, si)):
        torch.testing.assert_close(got_s, es, rtol=tol, atol=tol)
        same = got_i == ei
        assert same.float().mean() > 0.99
        assert bool(((got_s - es).abs()[~same] <= tol * es.abs()[~same] + 1e-6).all())     # a near-tie swapped
    if n <= 40_000:
        okapi = bm25_okapi.OkapiCsr(doc_off.cpu().numpy(), doc_tok.cpu().numpy(), vocab)
        terms = qb.q_terms.view(n_q, -1).cpu().numpy()
        bm = torch.tensor(np.stack([okapi.get_scores(terms[q]) for q in range(n_q)]), dtype=torch.float32)
        de = torch.tensor(dense_fusion.dense_scores(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy()),
                          dtype=torch.float32)
        state = {key: v.detach().cpu() for key, v in router.state_dict().items()}
        want_s, want_i = router_oracle.hybrid_rerank(bm, de, state, True, k)
        torch.testing.assert_close(ts.cpu(), want_s, rtol=3 * tol, atol=3 * tol)
        assert (ti.cpu().long() - 1000 == want_i).float().mean() > 0.98


@pytest.mark.parametrize("hidden,scale,dense_std", [(64, 1.0, 0.3), (32, 6.0, 0.05), (16, 20.0, 0.02), (128, 3.0, 0.1)])
def test_full_fusion_bound_as_the_kernel_indexes_it(rq, dev, hidden, scale, dense_std):
    """The gate-bound lookup EXACTLY as the fused epilogue evaluates it (FusedBound in csrc/dense_mma.cu, called
    through the debug entry point) on 600k (bm25, dense) pairs spanning every cell, with steep gates and a small
    dense_std: the bound never undercuts the fused score, the cell it reads is the floor cell (or a neighbour only
    for points within 1e-4 cells of a grid line), cosines close to +-d_hi do not wrap to the other end of the table."""
    import ctypes as C
    from rag_uq_b200 import _lib
    from rag_uq_b200.router import full_fusion_bounds
    fn = _lib.lib.ragb_debug_fused_bound
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                        C.c_float, C.c_void_p, C.c_void_p]
    torch.manual_seed(hidden)
    lin1, lin2 = torch.nn.Linear(3, hidden), torch.nn.Linear(hidden, 1)
    w1, b1 = (lin1.weight.detach() * scale).numpy(), (lin1.bias.detach() * scale).numpy()
    w2, b2 = (lin2.weight.detach() * scale).reshape(-1).numpy(), lin2.bias.detach().numpy()
    stats = np.array([8.0, 6.0, 0.2, dense_std], dtype=np.float32)
    b_cap, d_hi, n_b, n_d = 32.0, 1.015625, 128, 64
    table = full_fusion_bounds(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)
    lo = (table.view(np.uint32) << 16).view(np.float32)
    hi = (table.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    rng = np.random.default_rng(hidden)
    n = 600_000
    b = rng.uniform(0.0, b_cap * 1.05, n).astype(np.float32)
    d = rng.uniform(-d_hi, d_hi, n).astype(np.float32)
    b[:2000] = 0.0                                                    # documents without any query term
    d[2000:4000] = np.float32(d_hi) * rng.choice([-1.0, 1.0], 2000).astype(np.float32)     # the edges of the table
    d[4000:6000] = rng.uniform(0.98, 1.0, 2000).astype(np.float32)    # cos close to 1 must stay in the top cells
    tb, td = torch.tensor(b, device=dev), torch.tensor(d, device=dev)
    ttab = torch.tensor(table, device=dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.check(fn(tb.data_ptr(), td.data_ptr(), n, ttab.data_ptr(), n_b, n_d, b_cap, d_hi, out.data_ptr(),
                  torch.cuda.current_stream().cuda_stream))
    got = out.cpu().numpy()
    bt, dt = torch.tensor(b), torch.tensor(d)
    bn = (bt - stats[0]) / (torch.tensor(stats[1]) + 1e-6)
    dn = (dt - stats[2]) / (torch.tensor(stats[3]) + 1e-6)
    feats = torch.stack([bn, dn, dn - bn], -1)
    gate = torch.sigmoid(torch.relu(feats @ torch.tensor(w1).T + torch.tensor(b1)) @ torch.tensor(w2) + torch.tensor(b2))
    fused = (gate * dt + (1 - gate) * bt).numpy()
    assert (got >= fused - 1e-5 * np.abs(fused) - 1e-6).all(), "the kernel's bound undercuts a fused score"
    # which cell did the kernel read?  Recompute the bound for the floor cell and its neighbours in float64.
    xb, xd = b.astype(np.float64) * (n_b / b_cap), (d.astype(np.float64) + d_hi) * (n_d / (2 * d_hi))
    fb = np.floor(xb).astype(np.int64)
    fb = np.where((fb < 0) | (fb > n_b - 1) | (xb < 0.25), n_b - 1, fb)       # documented: tiny / huge bm25 -> last row (0, 1)
    fd = np.clip(np.floor(xd).astype(np.int64), 0, n_d - 1)

    def bound_at(ib, idx):
        g = np.where(d <= b, lo[ib, idx], hi[ib, idx]).astype(np.float32)
        return (g * (d - b) + b).astype(np.float32)

    # b + g (d - b) cancels when g is close to 1: the kernel's single-rounding fmaf and numpy's two roundings differ
    # by up to a few ulp of b (b <= 34: ulp 4e-6)
    same = np.isclose(got, bound_at(fb, fd), rtol=1e-6, atol=1.2e-5)
    near_line = (np.abs(xd - np.round(xd)) < 1e-4) | (np.abs(xb - np.round(xb)) < 1e-4) | (np.abs(xb - 0.25) < 1e-3)
    assert (same | near_line).all(), f"{int((~(same | near_line)).sum())} pairs read a cell that is not their floor cell"
    assert same.mean() > 0.999


@pytest.mark.parametrize("seed", list(range(6)))
def test_full_fusion_fused_epilogue_random_shapes(rq, dev, seed):
    """Differential test of the fused epilogue against the un-fused path on odd shapes: ragged tile tails, one and many
    query slabs, k from 1 to 100, hidden 16-128, gates of any steepness, statistics that put the gate anywhere."""
    from rag_uq_b200 import synth
    g = torch.Generator().manual_seed(77 + seed)
    n = int(torch.randint(300, 6000, (1,), generator=g))
    n_q = int(torch.randint(9, 300, (1,), generator=g))
    k = int(torch.randint(1, 101, (1,), generator=g))
    hidden = [16, 32, 64, 128, 64, 48][seed]
    scale = [1.0, 3.0, 0.3, 12.0, 1.0, 6.0][seed]
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, 768, dev)
    engine = rq.HybridEngine(rq.build_shard(doc_off, doc_tok, vocab).finalize(), passages, id_base=seed * 1000)
    qb = synth.make_queries(n_q, n, 768, cdf, dev)
    torch.manual_seed(seed)
    router = rq.RetrievalRouter(rq.RouterConfig(hidden_dim=hidden)).to(dev).eval()
    with torch.no_grad():
        for p in router.parameters():
            p.mul_(scale)
    stats = torch.rand(4, generator=g)
    router.bm25_mean.fill_(float(stats[0]) * 20 - 2); router.bm25_std.fill_(float(stats[1]) * 10 + 0.1)
    router.dense_mean.fill_(float(stats[2]) - 0.5); router.dense_std.fill_(float(stats[3]) * 0.5 + 0.02)
    router.stats_initialized = True
    with torch.no_grad():
        fs, fi = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=True)
        us, ui = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=False, query_chunk=64)
    assert fs.shape == us.shape == (n_q, min(k, n))
    tol = 1e-5 * max(1.0, scale * scale) * max(1.0, 0.2 / float(router.dense_std))
    torch.testing.assert_close(fs, us, rtol=tol, atol=tol)
    same = fi == ui
    assert same.float().mean() > 0.98
    assert bool(((fs - us).abs()[~same] <= tol * us.abs()[~same] + tol).all())


def test_row_sharded_engine_equals_single_engine(rq, dev):
    """Emulate G = 2 on one GPU: local pools per shard, merged exactly as the all-gather path merges."""
    from rag_uq_b200 import synth
    n, dim, k, pool, n_q = 9000, 768, 10, 50, 32
    vocab, cdf, doc_off, doc_tok = _synthetic_corpus(rq, dev, n)
    passages = synth.passage_embeddings(0, n, dim, dev)
    whole = rq.HybridEngine(rq.build_shard(doc_off, doc_tok, vocab).finalize(), passages)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    want = whole.hybrid_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, k, pool)
    shards = []
    for r in range(2):
        lo, hi = rq.shard_rows(n, 2, r)
        sp = rq.build_shard(doc_off[lo:hi + 1] - doc_off[lo], doc_tok[int(doc_off[lo]):int(doc_off[hi])], vocab, id_base=lo)
        shards.append((sp, passages[lo:hi].contiguous(), lo))
    df = shards[0][0].df + shards[1][0].df
    total = int(shards[0][0].doc_len.sum() + shards[1][0].doc_len.sum())
    pools_b, pools_d = [], []
    for sp, emb, lo in shards:
        sp.finalize(df, n, total)
        eng = rq.HybridEngine(sp, emb, id_base=lo)
        pools_b.append(sp.score_topk(qb.q_terms, qb.q_off, qb.max_terms, pool))
        pools_d.append(eng.dense_local_topk(qb.q_emb, pool))
    bs, bi = rq.ops.topk_merge(torch.stack([p[0] for p in pools_b], 1), torch.stack([p[1] for p in pools_b], 1), pool)
    ds, di = rq.ops.topk_merge(torch.stack([p[0] for p in pools_d], 1), torch.stack([p[1] for p in pools_d], 1), pool)
    got = rq.ops.hybrid_fuse_topk(bs, bi, ds, di, k)
    assert torch.equal(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-6)


# ------------------------------------------------------------------------------------------
# router + MC-Dropout
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "router_golden.npz")


def _router_from_golden(rq, gold, tag, dev):
    hidden = 64 if tag == "h64" else 32
    router = rq.RetrievalRouter(rq.RouterConfig(hidden_dim=hidden))
    pre = f"{tag}/state/"
    router.load_state_dict({k[len(pre):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(pre)})
    return router.to(dev).eval()


@pytest.mark.parametrize("tag", ["h64", "h32"])
def test_router_matches_reference_golden(rq, dev, gold, tag):
    router = _router_from_golden(rq, gold, tag, dev)
    b, d = torch.from_numpy(gold[f"{tag}/bm25"]).to(dev), torch.from_numpy(gold[f"{tag}/dense"]).to(dev)
    with torch.no_grad():
        g0 = router(b, d)                                              # stats_initialized False after load
        np.testing.assert_allclose(g0.cpu().numpy(), gold[f"{tag}/gate_batchstat"], rtol=1e-5, atol=1e-6)
        vals, idx = router.hybrid_rerank(b, d, top_k=10)
        assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), gold[f"{tag}/rerank_batchstat_idx"])
        np.testing.assert_allclose(vals.cpu().numpy(), gold[f"{tag}/rerank_batchstat_vals"], rtol=1e-5, atol=1e-5)
        for name, val in zip(["bm25_mean", "bm25_std", "dense_mean", "dense_std"], gold[f"{tag}/running_stats"]):
            getattr(router, name).fill_(float(val))
        router.stats_initialized = True
        g1 = router(b, d)
        np.testing.assert_allclose(g1.cpu().numpy(), gold[f"{tag}/gate_running"], rtol=1e-5, atol=1e-6)
        vals, idx = router.hybrid_rerank(b, d, top_k=10)
        assert np.array_equal(idx.cpu().numpy(), gold[f"{tag}/rerank_running_idx"])
        bb, bd = torch.from_numpy(gold[f"{tag}/big_bm25"]).to(dev), torch.from_numpy(gold[f"{tag}/big_dense"]).to(dev)
        vals, idx = router.hybrid_rerank(bb, bd, top_k=500)           # k > P clamps (router.py:202); P=300 > 256 ...
    assert vals.shape == (2, 300)


def test_router_reference_test_suite_properties(rq, dev):
    """The reference's own router tests (tests/test_router.py:43-131), run against the drop-in."""
    torch.manual_seed(0)
    router = rq.RetrievalRouter().to(dev).eval()
    with torch.no_grad():
        w = router(torch.randn(4, 20, device=dev), torch.randn(4, 20, device=dev))
        assert w.shape == (4, 20) and (w >= 0).all() and (w <= 1).all()
        assert router(torch.randn(1, 10, device=dev), torch.randn(1, 10, device=dev)).shape == (1, 10)
        b = torch.tensor([[1.0, 2.0, 3.0, 4.0, 5.0]], device=dev)
        d = torch.tensor([[5.0, 4.0, 3.0, 2.0, 1.0]], device=dev)
        s, i = router.hybrid_rerank(b, d, top_k=3)
        assert s.shape == (1, 3) and i.shape == (1, 3)
        s, i = router.hybrid_rerank(torch.randn(2, 5, device=dev), torch.randn(2, 5, device=dev), top_k=10)
        assert s.shape == (2, 5)
        dec = router.get_routing_decision(torch.randn(2, 10, device=dev), torch.randn(2, 10, device=dev))
        assert {"avg_dense_weight", "weight_std", "dense_preferred_ratio", "bm25_preferred_ratio", "routing_weights"} <= set(dec)
        assert 0 <= dec["avg_dense_weight"] <= 1 and dec["routing_weights"].shape == (2, 10)
        assert not router.stats_initialized
        router.train()
        router(torch.randn(4, 20, device=dev) * 10, torch.randn(4, 20, device=dev))
        assert router.stats_initialized                                  # tests/test_router.py:108-119
    with pytest.raises(NotImplementedError):
        router(torch.randn(2, 3), torch.randn(2, 3))                     # CPU tensors: no fallback
    assert torch.isnan(rq.RetrievalRouter().to(dev).eval()(torch.ones(1, 1, device=dev), torch.ones(1, 1, device=dev))).all()


def test_router_large_and_bf16_inputs(rq, dev):
    """[64, 10000] candidates (full-fusion shape of config C1); fp32 1e-5, bf16-rounded inputs 1e-3."""
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    state = {k: v.detach().cpu() for k, v in router.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    b, d = torch.rand(64, 10_000, generator=g) * 25, torch.rand(64, 10_000, generator=g) * 2 - 1
    b[b < 12] = 0.0                                                     # BM25 is sparse
    with torch.no_grad():
        for armed in (False, True):
            if armed:
                router.bm25_mean.fill_(3.0); router.bm25_std.fill_(6.0)
                router.dense_mean.fill_(0.0); router.dense_std.fill_(0.6)
                router.stats_initialized = True
                state = {k: v.detach().cpu() for k, v in router.state_dict().items()}
            got = router(b.to(dev), d.to(dev)).cpu()
            want = router_oracle.gate(b, d, state, armed)
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
            vals, idx = router.hybrid_rerank(b.to(dev), d.to(dev), top_k=10)
            ovals, oidx = router_oracle.hybrid_rerank(b, d, state, armed, 10)
            torch.testing.assert_close(vals.cpu(), ovals, rtol=1e-5, atol=1e-5)
            assert (idx.cpu() == oidx).float().mean() > 0.98
        got16 = router(b.to(dev).bfloat16(), d.to(dev).bfloat16()).cpu()
        want16 = router_oracle.gate(b.bfloat16().float(), d.bfloat16().float(), state, True)
        torch.testing.assert_close(got16, want16, rtol=1e-3, atol=1e-3)
        # per-query statistics == calling the reference once per query (run_evaluation.py:171-177)
        router.stats_initialized = False
        rows = router(b[:5, :10].contiguous().to(dev), d[:5, :10].contiguous().to(dev), per_query_stats=True).cpu()
        for q in range(5):
            want = router_oracle.gate(b[q:q + 1, :10], d[q:q + 1, :10], state, False)
            torch.testing.assert_close(rows[q:q + 1], want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag,layout", [("h64", 0), ("h64", 1), ("h32", 0), ("h32", 1)])
def test_mc_dropout_masks_are_philox_and_gates_match_injected_masks(rq, dev, gold, tag, layout):
    router = _router_from_golden(rq, gold, tag, dev)
    hidden = router.config.hidden_dim
    b, d = torch.from_numpy(gold[f"{tag}/bm25"]).to(dev), torch.from_numpy(gold[f"{tag}/dense"]).to(dev)
    T, seed, offset = 5, 0xC0FFEE, 8
    with torch.no_grad():
        unc = router.mc_dropout(b, d, n_samples=T, seed=seed, offset=offset, torch_layout=bool(layout), return_samples=True)
    masks = unc.masks.cpu().numpy()                                    # [T, B*P, H]
    n_el = b.numel() * hidden
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    for t in range(T):
        if layout == 1:
            inc = philox.torch_dropout_geometry(n_el, sm)[2]
            want = philox.keep_mask_torch_layout(n_el, seed, offset + t * inc, 0.9, sm)
        else:
            cand = np.repeat(np.arange(b.numel(), dtype=np.uint64), hidden // 4)
            quad = np.tile(np.arange(hidden // 4, dtype=np.uint64), b.numel())
            bits = philox.draw4(seed, cand, np.uint64(offset // 4) + np.uint64(t * (hidden // 4)) + quad)
            want = (philox.uniform_from_bits(bits).reshape(-1) < np.float32(0.9)).astype(np.uint8)
        assert np.array_equal(masks[t].reshape(-1), want), (t, layout)            # bit-exact
    state = {k: v.detach().cpu() for k, v in router.state_dict().items()}
    ref = router_oracle.mc_dropout(b.cpu(), d.cpu(), state, False, torch.from_numpy(masks).float())
    for t in range(T):
        want = router_oracle.gate(b.cpu(), d.cpu(), state, False, keep_mask=torch.from_numpy(masks[t]).float())
        torch.testing.assert_close(unc.gates[t].cpu(), want, rtol=1e-5, atol=1e-5)  # mask-injection parity
    torch.testing.assert_close(unc.mean_gate.cpu(), ref["mean_w"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(unc.std_gate.cpu(), ref["std_w"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(unc.mean_fused.cpu(), ref["mean_h"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(unc.std_fused.cpu(), ref["std_h"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(unc.variance.cpu(), ref["variance"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(unc.confidence.cpu(), ref["confidence"], rtol=1e-4, atol=1e-6)
    assert torch.equal(unc.consensus.cpu(), ref["consensus"])
    res = unc.to_confidence_result(0)
    assert res.confidence == pytest.approx(1 - min(1.0, res.embedding_variance / 2))


def test_mc_dropout_reference_golden_masks_statistics(rq, dev, gold):
    """Masks from torch-CPU dropout (golden, from the live reference) cannot be reproduced in-kernel
    (mt19937); check instead that T = 200 Philox samples have the same mean gate within sampling error."""
    router = _router_from_golden(rq, gold, "h64", dev)
    b, d = torch.from_numpy(gold["h64/bm25"]).to(dev), torch.from_numpy(gold["h64/dense"]).to(dev)
    for name, val in zip(["bm25_mean", "bm25_std", "dense_mean", "dense_std"], gold["h64/running_stats"]):
        getattr(router, name).fill_(float(val))
    router.stats_initialized = True
    with torch.no_grad():
        unc = router.mc_dropout(b, d, n_samples=200, seed=1)
    ref_mean = gold["h64/mc_gates"].mean(axis=0)
    ref_sem = gold["h64/mc_gates"].std(axis=0) / math.sqrt(gold["h64/mc_gates"].shape[0]) + 1e-3
    assert np.all(np.abs(unc.mean_gate.cpu().numpy() - ref_mean) < 6 * ref_sem)


def test_mc_dropout_torch_cuda_bit_exact(rq, dev):
    """Stretch goal of SURVEY H6: sample t == the t-th train-mode forward of a torch router on this GPU."""
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    lin1, lin2 = router.scorer[0], router.scorer[3]
    g = torch.Generator().manual_seed(4)
    b, d = (torch.rand(16, 100, generator=g) * 10).to(dev), torch.rand(16, 100, generator=g).to(dev)
    T = 3
    torch.cuda.manual_seed(2024)
    gen = torch.cuda.default_generators[0]
    seed, offset0 = gen.initial_seed(), gen.get_offset()
    with torch.no_grad():
        bn = (b - b.mean()) / (b.std() + 1e-6)
        dn = (d - d.mean()) / (d.std() + 1e-6)
        feats = torch.stack([bn, dn, dn - bn], -1).view(-1, 3)
        hidden = torch.relu(torch.nn.functional.linear(feats, lin1.weight, lin1.bias))
        torch_gates = []
        for _ in range(T):
            h = torch.nn.functional.dropout(hidden, 0.1, training=True)
            torch_gates.append(torch.sigmoid(torch.nn.functional.linear(h, lin2.weight, lin2.bias)).view(16, 100))
        consumed = gen.get_offset() - offset0
        unc = router.mc_dropout(b, d, n_samples=T, seed=seed, offset=offset0, torch_layout=True, return_samples=True)
    from rag_uq_b200.router import torch_dropout_increment
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    assert consumed == T * torch_dropout_increment(hidden.numel(), sm)
    for t in range(T):
        torch.testing.assert_close(unc.gates[t], torch_gates[t], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------
# drop-in string API
# ------------------------------------------------------------------------------------------
def _pool_tensors(pool_rows, number, width, dev):
    score = torch.zeros((1, width), dtype=torch.float32)
    ident = torch.full((1, width), -1, dtype=torch.int32)
    for j, (doc_id, sc) in enumerate(pool_rows[:width]):
        score[0, j], ident[0, j] = sc, number(doc_id)
    return score.to(dev), ident.to(dev)


def test_hybrid_fuse_kernel_vs_live_reference_golden(rq, dev, retrieval_gold):
    """ragb_hybrid_fuse_topk against what the reference's own HybridRetriever.hybrid_search returned
    (streaming_index.py:464-523, run live with stubbed pools by tests/golden/make_retrieval_golden.py)."""
    hy = retrieval_gold["hybrid"]
    known = {d: i for i, d in enumerate(hy["document_ids"])}
    for case in hy["cases"]:
        pool = case["retrieval_pool_size"]
        width = max(pool, 1)
        # ids the retriever holds no document for never reach the fusion (:494-496): id -1 / score 0 = "no entry"
        bs, bi = _pool_tensors([r for r in case["bm25_pool"][:pool] if r[0] in known], known.get, width, dev)
        ds, di = _pool_tensors([r for r in case["dense_pool"][:pool] if r[0] in known], known.get, width, dev)
        k = min(case["top_k"], 2 * width)
        ids, ob, od, oh = (t[0].cpu().tolist() for t in rq.ops.hybrid_fuse_topk(bs, bi, ds, di, k))
        want = case["hybrid_search"]
        n = len(want)
        assert [i for i in ids if i >= 0] == ids[:n] and all(i == -1 for i in ids[n:]), case["name"]
        assert_ranking_matches(ids[:n], oh[:n], [known[w[0]] for w in want], [w[3] for w in want], None, 2e-6, 1e-7)
        by_id = {known[w[0]]: w for w in want}
        for i, b, d in zip(ids[:n], ob[:n], od[:n]):
            assert b == np.float32(by_id[i][1]) and d == np.float32(by_id[i][2]), case["name"]   # scores pass through untouched


def test_hybrid_retriever_dropin_vs_live_reference_golden(rq, dev, retrieval_gold):
    """The drop-in HybridRetriever.hybrid_search / get_scores_for_router (host join on document ids + fusion kernel)
    fed the SAME pools the live reference was fed: same documents, scores, order (modulo ties), texts, titles, padding."""
    hy = retrieval_gold["hybrid"]
    names = hy["document_ids"]
    for case in hy["cases"]:
        r = rq.HybridRetriever(bm25_persist_path=None, chroma_persist_path=None)
        for i, d in enumerate(names):
            r.documents[d] = rq.Document(id=d, text=f"text of {d}", title=f"title {d}")
            r._order[d] = i
        # the two indices number their rows independently and may hold ids the retriever has no document for
        b_rows = names + sorted({x[0] for x in case["bm25_pool"]} - set(names))
        d_rows = list(reversed(names)) + sorted({x[0] for x in case["dense_pool"]} - set(names))
        r.bm25_index.doc_ids, r.dense_index.ids = b_rows, d_rows

        def stub(queries, query_embeddings=None, top_k=10, retrieval_pool_size=50, case=case, b_rows=b_rows, d_rows=d_rows):
            w = max(retrieval_pool_size, 1)
            bm = _pool_tensors(case["bm25_pool"], b_rows.index, w, dev) if case["bm25_pool"] else None
            de = _pool_tensors(case["dense_pool"], d_rows.index, w, dev) if case["dense_pool"] else None
            return 1, bm, de

        r.hybrid_search_batch = stub
        got = r.hybrid_search("ignored", top_k=case["top_k"], retrieval_pool_size=case["retrieval_pool_size"])
        want = case["hybrid_search"]
        assert_ranking_matches([g.doc_id for g in got], [g.hybrid_score for g in got], [w[0] for w in want],
                               [w[3] for w in want], None, 2e-6, 1e-7)
        by_id = {w[0]: w for w in want}
        for g in got:
            w = by_id[g.doc_id]
            assert (g.bm25_score, g.dense_score, g.text, g.title) == (float(np.float32(w[1])), float(np.float32(w[2])), w[4], w[5])
        ref = case["scores_for_router"]
        bsc, dsc, ids, txt = r.get_scores_for_router("ignored", num_passages=case["num_passages"])
        n_real = sum(1 for i in ref["ids"] if i)
        assert len(ids) == len(ref["ids"]) and ids[n_real:] == [""] * (len(ids) - n_real) and txt[n_real:] == ref["texts"][n_real:]
        assert bsc[n_real:] == [0.0] * (len(ids) - n_real) and dsc[n_real:] == [0.0] * (len(ids) - n_real)
        assert sorted(ids[:n_real]) == sorted(ref["ids"][:n_real]), case["name"]
        ref_rows = {i: (b, d, t) for i, b, d, t in zip(ref["ids"], ref["bm25"], ref["dense"], ref["texts"])}
        for i, b, d, t in zip(ids[:n_real], bsc, dsc, txt):
            assert (b, d, t) == (float(np.float32(ref_rows[i][0])), float(np.float32(ref_rows[i][1])), ref_rows[i][2])


def test_bm25_index_search_vs_live_reference_golden(rq, dev, retrieval_gold):
    """The drop-in BM25Index.search (host tokeniser + vocabulary + CSR segments + bm25_kernel) against the live
    reference's BM25Index.search (streaming_index.py:150-179) on the same texts: same documents in the same order
    (modulo ties), scores within 1e-5 relative, nothing with score <= 0, top_k > N."""
    for corpus in retrieval_gold["bm25_index_search"]:
        index = rq.BM25Index(k1=corpus["k1"], b=corpus["b"])
        docs = [rq.Document(id=d, text=t) for d, t in zip(corpus["doc_ids"], corpus["texts"])]
        half = len(docs) // 2 + 1
        assert index.add_documents(docs[:half]) == half and index.add_documents(docs) == len(docs) - half
        lit = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in corpus["texts"]], k1=corpus["k1"], b=corpus["b"])
        for case in corpus["cases"]:
            got = index.search(case["query"], case["top_k"])
            want = case["result"]
            full = dict(zip(corpus["doc_ids"], lit.get_scores(bm25_okapi.tokenize(case["query"])).tolist()))
            assert_ranking_matches([g[0] for g in got], [g[1] for g in got], [w[0] for w in want], [w[1] for w in want],
                                   full, 1e-5, 1e-9)
            assert all(s > 0 for _, s in got)


def test_hybrid_retriever_dropin(rq, dev, tmp_path):
    rng = np.random.default_rng(0)
    words = [f"w{i}" for i in range(200)]
    texts = [" ".join(rng.choice(words, size=rng.integers(5, 30))) for _ in range(300)]
    table = {t: rng.standard_normal(96).astype(np.float32) for t in texts}

    def embed(batch):
        return np.stack([table.get(t, np.ones(96, np.float32)) for t in batch])

    r = rq.HybridRetriever(bm25_persist_path=str(tmp_path / "bm25.pkl"), chroma_persist_path=str(tmp_path / "dense"),
                           embed_fn=embed)
    docs = [rq.Document(id=f"doc{i}", text=t, title=f"T{i}") for i, t in enumerate(texts)]
    stats = r.add_documents(docs[:200])
    assert stats == {"bm25_added": 200, "dense_added": 200, "total_documents": 200}
    stats = r.add_documents(docs[150:])
    assert stats == {"bm25_added": 100, "dense_added": 100, "total_documents": 300} and len(r) == 300

    okapi = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in texts])
    emb = r.dense_index.matrix.float().cpu().numpy()          # the stored bf16 rows are the ground truth
    for query in [texts[17], "w3 w3 w77 unknownword", texts[250][:40]]:
        qe = r.dense_index._to_rows(embed([query])).float().cpu().numpy()
        bm = bm25_okapi.index_search(okapi.get_scores(bm25_okapi.tokenize(query)), 50)
        de = dense_fusion.topk_desc(dense_fusion.dense_scores(emb, qe), 50)[0]
        want = dense_fusion.hybrid_search(bm, de, 10)
        got = r.hybrid_search(query, top_k=10)
        assert [g.doc_id for g in got] == [f"doc{w[0]}" for w in want]
        for g_, w in zip(got, want):
            assert g_.bm25_score == pytest.approx(w[1], rel=1e-5, abs=1e-7)
            assert g_.dense_score == pytest.approx(w[2], abs=3e-6)
            assert g_.hybrid_score == pytest.approx(w[3], rel=2e-5, abs=1e-6)
            assert g_.text == texts[w[0]] and g_.title == f"T{w[0]}"
        assert [d for d, _ in r.bm25_search(query, 20)] == [f"doc{i}" for i, _ in bm[:20]]
        for num in (20, 7, 120):
            bsc, dsc, ids, txt = r.get_scores_for_router(query, num_passages=num)
            ob, od, oi = dense_fusion.scores_for_router(bm, de, num)          # pools of 50, as :537 implies
            assert len(bsc) == len(dsc) == len(ids) == len(txt) == num
            assert ids == [f"doc{i}" if i >= 0 else "" for i in oi]
            assert txt == [texts[i] if i >= 0 else "" for i in oi]
            np.testing.assert_allclose(bsc, ob, rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(dsc, od, rtol=0, atol=3e-6)
    # persistence uses the reference's pickle schema and reloads
    again = rq.BM25Index(persist_path=str(tmp_path / "bm25.pkl"))
    assert len(again) == 300 and again.search(texts[17], 5) == r.bm25_index.search(texts[17], 5)
    import pickle
    with open(tmp_path / "bm25.pkl", "rb") as fh:
        assert set(pickle.load(fh)) == {"documents", "doc_ids", "tokenized_corpus", "k1", "b"}
    # a restarted process gets BOTH sides back (the dense rows are persisted next to the BM25 pickle) and answers
    # exactly as before; a StreamingIndex checkpoint taken then may be resumed without losing dense rows
    before = r.hybrid_search(texts[17], top_k=10)
    r2 = rq.HybridRetriever(bm25_persist_path=str(tmp_path / "bm25.pkl"), chroma_persist_path=str(tmp_path / "dense"),
                            embed_fn=embed)
    assert len(r2) == 300 and len(r2.dense_index) == 300 and len(r2.bm25_index) == 300
    assert torch.equal(r2.dense_index.matrix.view(torch.int16), r.dense_index.matrix.view(torch.int16))
    assert r2.dense_index.ids == r.dense_index.ids and r2.dense_index.texts == r.dense_index.texts
    after = r2.hybrid_search(texts[17], top_k=10)
    assert [(a.doc_id, a.bm25_score, a.dense_score, a.hybrid_score, a.text) for a in after] == \
        [(b.doc_id, b.bm25_score, b.dense_score, b.hybrid_score, b.text) for b in before]
    assert r2.add_documents(docs[:50]) == {"bm25_added": 0, "dense_added": 0, "total_documents": 300}
    # without a dense store the resume guard refuses to skip the checkpointed lines
    corpus = tmp_path / "c.jsonl"
    corpus.write_text("".join(json.dumps({"id": f"s{i}", "text": texts[i]}) + "\n" for i in range(6)))
    r3 = rq.HybridRetriever(bm25_persist_path=str(tmp_path / "b3.pkl"), chroma_persist_path=None, embed_fn=embed)
    assert list(rq.StreamingIndex(r3, str(tmp_path / "ck3.json"), batch_size=4).stream_from_jsonl(str(corpus))) == [4, 2]
    r4 = rq.HybridRetriever(bm25_persist_path=str(tmp_path / "b3.pkl"), chroma_persist_path=None, embed_fn=embed)
    resumed = rq.StreamingIndex(r4, str(tmp_path / "ck3.json"), batch_size=4)
    assert resumed.progress["last_offset"] == 0                         # checkpoint ignored: dense rows were not persisted
    assert list(resumed.stream_from_jsonl(str(corpus))) == [4, 2] and len(r4.dense_index) == 6 and len(r4.bm25_index) == 6


def test_incremental_ingest_equals_full_rebuild(rq, dev):
    """N1: documents added in batches (one GPU segment per batch, statistics refreshed, segments merged
    beyond 8) score exactly like the reference semantics, which rebuild BM25Okapi over everything on every add."""
    rng = np.random.default_rng(3)
    words = [f"t{i}" for i in range(400)]
    probs = 1.0 / np.arange(1, 401)
    probs /= probs.sum()
    texts = [" ".join(rng.choice(words, size=rng.integers(8, 60), p=probs)) for _ in range(1300)]
    docs = [rq.Document(id=f"d{i}", text=t) for i, t in enumerate(texts)]
    index = rq.BM25Index()
    sizes = [1, 7, 300, 5, 120, 64, 33, 250, 11, 200, 309]          # 11 batches -> exercises the merge
    queries = ["t0 t1 t17 t250", "t399 t3 t3", "t5", "unknown t2 t390 t391"]
    done = 0
    for n in sizes:
        index.add_documents(docs[done:done + n])
        done += n
        okapi = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in texts[:done]])
        for q in queries:
            want = bm25_okapi.index_search(okapi.get_scores(bm25_okapi.tokenize(q)), 20)
            got = index.search(q, top_k=20)
            assert len(got) == len(want), (done, q)
            for (gid, gs), (wi, ws) in zip(got, want):
                assert gs == pytest.approx(ws, rel=1e-5), (done, q)
                assert gid == f"d{wi}" or okapi.get_scores(bm25_okapi.tokenize(q))[int(gid[1:])] == pytest.approx(ws, rel=2e-6)
        assert len(index.bm25.segments) <= 8
    assert done == 1300 and index.bm25.n_docs == 1300
    fresh = rq.BM25Index()
    fresh.add_documents(docs)
    for q in queries:
        a, b = index.search(q, 50), fresh.search(q, 50)
        assert [d for d, _ in a] == [d for d, _ in b]
        np.testing.assert_allclose([s for _, s in a], [s for _, s in b], rtol=2e-6)


def test_shard_directory_reload_gives_identical_results(rq, dev, tmp_path):
    """N2: save the engine's shard as raw arrays, load it back (idf / norm / table rows recomputed on the device)
    and get bit-identical hybrid results."""
    from rag_uq_b200 import synth
    n, dim, n_q = 20_000, 768, 40
    engine, cdf = synth.build_synthetic_engine(n, dim, dev)
    qb = synth.make_queries(n_q, n, dim, cdf, dev)
    want = engine.hybrid_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, 10, 50)
    rq.save_engine(tmp_path / "shard", engine)
    again = rq.load_engine(tmp_path / "shard", dev)
    assert torch.equal(again.sparse.idf, engine.sparse.idf) and torch.equal(again.sparse.norm, engine.sparse.norm)
    assert torch.equal(again.sparse.dense_terms, engine.sparse.dense_terms)
    got = again.hybrid_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, 10, 50)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_graphed_small_batch_search_equals_eager(rq, dev):
    """C2 latency path: the whole batch-1 (and batch-4) step captured as one CUDA graph - BM25 chain and GEMV as parallel
    branches - returns exactly what the eager call sequence returns, replay after replay, also for ragged queries."""
    from rag_uq_b200 import synth
    n, dim = 60_000, 768
    engine, cdf = synth.build_synthetic_engine(n, dim, dev)
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.2); router.dense_std.fill_(0.3)
    router.stats_initialized = True
    for batch in (1, 4):
        gs = engine.graphed_search(router, batch, max_terms=8, k=10, pool=50)
        for first in (0, 17, 400):
            qb = synth.make_queries(batch, n, dim, cdf, dev, first_query=first)
            with torch.no_grad():
                want = engine.retrieve_and_rerank(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, 10, 50)
                ids, fused, sb, sd = (t.clone() for t in gs(qb.q_terms, qb.q_off, qb.q_emb))
            assert torch.equal(ids, want["ids"]) and torch.equal(fused, want["fused"])
            assert torch.equal(sb, want["bm25"]) and torch.equal(sd, want["dense"])
        # a ragged batch (3 and 5 tokens ...) goes through the padded static buffers
        qb = synth.make_queries(batch, n, dim, cdf, dev, first_query=900)
        lens = [3 + (i % 5) for i in range(batch)]
        rows = qb.q_terms.view(batch, -1)
        flat = torch.cat([rows[i, :lens[i]] for i in range(batch)]).contiguous()
        off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
        with torch.no_grad():
            want = engine.retrieve_and_rerank(flat, off, 8, qb.q_emb, router, 10, 50)
            ids, fused, _, _ = gs(flat, off, qb.q_emb)
        assert torch.equal(ids, want["ids"]) and torch.equal(fused, want["fused"])
    with pytest.raises(ValueError):
        engine.graphed_search(router, 9, 8)


def test_batched_string_api_tokenizer_and_pool_join(rq, dev, tmp_path):
    """N3: hybrid_search_many = one tokeniser / vocabulary pass and one device-side pool join for the whole batch; it
    returns exactly what per-query hybrid_search returns (also when the retriever holds no document for some rows)."""
    rng = np.random.default_rng(4)
    words = [f"w{i}" for i in range(300)]
    texts = [" ".join(rng.choice(words, size=rng.integers(5, 40))) for _ in range(500)]
    table = {t: rng.standard_normal(64).astype(np.float32) for t in texts}
    embed = lambda batch: np.stack([table.get(t, np.ones(64, np.float32)) for t in batch])   # noqa: E731
    r = rq.HybridRetriever(bm25_persist_path=None, chroma_persist_path=None, embed_fn=embed)
    r.add_documents([rq.Document(id=f"doc{i}", text=t) for i, t in enumerate(texts)])
    queries = [texts[3], "w1 W2  w2 unknown", "", texts[77][:25], "zzz qqq"] + [texts[i] for i in range(100, 130)]
    terms, off, longest = r.bm25_index.encode_queries(queries)
    want_rows = [[r.bm25_index.vocab.get(t, -1) for t in q.lower().split()] for q in queries]
    assert off.cpu().tolist() == np.concatenate([[0], np.cumsum([len(x) for x in want_rows])]).tolist()
    assert terms.cpu().tolist()[:int(off[-1])] == [t for row in want_rows for t in row] and longest == max(len(x) for x in want_rows)
    many = r.hybrid_search_many(queries, None, top_k=10, retrieval_pool_size=50)
    for q, got in zip(queries, many):
        one = r.hybrid_search(q, top_k=10, retrieval_pool_size=50)
        # the batch goes through the tcgen05 kernel, a single query through the GEMV: same dot products, another
        # summation order (1 ulp), BM25 bit-identical
        assert [(g.doc_id, g.bm25_score) for g in got] == [(g.doc_id, g.bm25_score) for g in one]
        np.testing.assert_allclose([g.dense_score for g in got], [g.dense_score for g in one], rtol=0, atol=1e-6)
        np.testing.assert_allclose([g.hybrid_score for g in got], [g.hybrid_score for g in one], rtol=0, atol=1e-6)
    # rows the retriever holds no document for are dropped before the fusion (:494-496): forget 100 documents
    for i in range(0, 500, 5):
        del r.documents[f"doc{i}"]
    r.__dict__.pop("_row_number_cache", None)
    after = r.hybrid_search_many(queries[:8], None, top_k=10)
    gone = {f"doc{i}" for i in range(0, 500, 5)}
    assert all(g.doc_id not in gone for res in after for g in res) and any(len(res) for res in after)


def test_retrieval_uncertainty(rq, dev):
    g = torch.Generator().manual_seed(9)
    scores = torch.rand(7, 10, generator=g)
    ids = torch.arange(70, dtype=torch.int32).view(7, 10)
    ids[2, 6:] = -1
    ids[5, :] = -1
    got = rq.ops.retrieval_uncertainty(scores.to(dev), ids.to(dev), 0.5).cpu()
    for q in range(7):
        valid = [float(s) for s, i in zip(scores[q], ids[q]) if i >= 0]
        assert float(got[q]) == pytest.approx(dense_fusion.retrieval_uncertainty(valid, 0.5), rel=1e-5, abs=1e-6)


def _check_against_streamed_oracle(rq, dev, engine, cdf, n, first_query, n_q, k, pool, router, tag):
    """Hybrid top-k + router rerank of n_q queries against oracle/large_check.py streamed over the WHOLE corpus."""
    from oracle import large_check
    from rag_uq_b200 import synth
    qb = synth.make_queries(n_q, n, 768, cdf, dev, first_query=first_query)
    with torch.no_grad():
        res = engine.retrieve_and_rerank(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, pool)
    state = {key: v.detach().cpu() for key, v in router.state_dict().items()}
    recs, info = large_check.run_synthetic_check(synth, dev, n, 768, qb.q_terms.view(n_q, -1).cpu().tolist(), qb.q_emb, pool, k,
                                                 state, True, df_expect=engine.sparse.df_global)
    assert info["df_matches_product"] is True
    ids, vals = res["ids"].cpu().tolist(), res["fused"].cpu().tolist()
    exact = 0
    for q, rec in enumerate(recs):
        ex, ok = large_check.compare_ranking(ids[q], vals[q], rec["rerank_ids"], rec["rerank_vals"], rec["rerank_all"], 2e-5, 1e-6)
        # a pool cut that fp32 and fp64 may place differently changes which documents are fused at all: not a kernel error
        assert ok or min(rec["bm25_gap"], rec["dense_gap"]) < 1e-5, (tag, q, ids[q], rec["rerank_ids"])
        exact += ex
    assert exact >= n_q - 1, (tag, exact)
    return exact


def test_c2_shape_1m_batch1_against_streamed_oracle(rq, dev):
    """BASELINE configs[1] at full size: 1M passages, batch-1 queries through the GEMV + BM25 path, top-10, checked
    against the float64 oracle over all 1M passages (ids identical modulo proven ties, scores 2e-5); then the same
    corpus through the batched tcgen05 path (16 queries)."""
    from rag_uq_b200 import synth
    n = 1_000_000
    engine, cdf = synth.build_synthetic_engine(n, 768, dev)
    torch.manual_seed(7)
    router = rq.RetrievalRouter().to(dev).eval()
    router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.2); router.dense_std.fill_(0.3)
    router.stats_initialized = True
    for first in (0, 1, 2):                               # three separate batch-1 calls (GEMV path)
        _check_against_streamed_oracle(rq, dev, engine, cdf, n, first, 1, 10, 50, router, f"b1-{first}")
    _check_against_streamed_oracle(rq, dev, engine, cdf, n, 100, 16, 10, 50, router, "b16-mma")


@pytest.mark.parametrize("n,n_q", [(2_000_000, 256)])
def test_large_corpus_properties(rq, dev, n, n_q):
    """Size-independent properties at a corpus the CPU oracle cannot score: the pruned / seeded top-k
    kernels against the library's own exhaustive score kernels, sortedness, uniqueness, shard invariance."""
    from rag_uq_b200 import synth
    k = 50
    engine, cdf = synth.build_synthetic_engine(n, 768, dev)
    qb = synth.make_queries(n_q, n, 768, cdf, dev)
    bs, bi = engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k)
    ds, di = engine.dense_local_topk(qb.q_emb, k)
    for score, ids in ((bs, bi), (ds, di)):
        assert int(ids.min()) >= 0 and int(ids.max()) < n                          # full lists, valid rows
        assert bool((score[:, :-1] >= score[:, 1:]).all())                          # sorted best first
        tie = score[:, :-1] == score[:, 1:]
        assert bool((ids[:, :-1][tie] < ids[:, 1:][tie]).all())                     # ties: lower id first
        assert all(len(set(r)) == k for r in ids[:8].tolist())                      # no duplicates
    # exhaustive check on a few queries: top-k of the full score vectors (library sort only as the checker)
    sub = slice(0, 6)
    t0, t1 = int(qb.q_off[0]), int(qb.q_off[6])
    full_b = engine.sparse.scores(qb.q_terms[t0:t1].contiguous(), (qb.q_off[0:7] - t0).contiguous(), qb.max_terms)
    full_d = rq.ops.dense_scores(engine.passages, qb.q_emb[sub].contiguous())
    for full, score, ids, tol in ((full_b, bs, bi, 1e-6), (full_d, ds, di, 2e-6)):
        want_s, want_i = torch.topk(full, k, dim=1)
        torch.testing.assert_close(score[sub], want_s, rtol=tol, atol=tol)
        got_at_ids = torch.gather(full, 1, ids[sub].long())
        torch.testing.assert_close(score[sub], got_at_ids, rtol=tol, atol=tol)      # reported score == score of that row
        assert (ids[sub].long() == want_i).float().mean() > 0.97                    # identical modulo ties
    # the source passage is the top dense hit, and its BM25 score is among the query's best
    assert (di[:, 0].long() == qb.source_rows).float().mean() > 0.95
    # shard invariance at scale: two row shards with global statistics, merged == unsharded, bit for bit
    lo, hi = rq.shard_rows(n, 2, 1)
    vocab = engine.sparse.vocab
    parts_b, parts_d = [], []
    shards = []
    for r in range(2):
        a, b = rq.shard_rows(n, 2, r)
        off, tok = synth.doc_tokens(a, b, cdf)
        shards.append((rq.build_shard(off, tok, vocab, id_base=a), a, b))
    df = shards[0][0].df + shards[1][0].df
    total = int(shards[0][0].doc_len.sum() + shards[1][0].doc_len.sum())
    for sp, a, b in shards:
        sp.finalize(df, n, total)
        parts_b.append(sp.score_topk(qb.q_terms, qb.q_off, qb.max_terms, k))
        parts_d.append(rq.ops.dense_mma_topk(engine.passages[a:b].contiguous(), qb.q_emb, k, a, 2))
    ms, mi = rq.ops.topk_merge(torch.stack([p[0] for p in parts_b], 1), torch.stack([p[1] for p in parts_b], 1), k)
    assert torch.equal(mi, bi) and torch.equal(ms, bs)
    ms, mi = rq.ops.topk_merge(torch.stack([p[0] for p in parts_d], 1), torch.stack([p[1] for p in parts_d], 1), k)
    assert torch.equal(mi, di) and torch.equal(ms, ds)


def test_error_behaviour(rq, dev):
    with pytest.raises(ValueError):
        rq.ops.topk_rows(torch.zeros(2, 10000, device=dev), 1000)           # k beyond RAGB_MAX_TOPK on a long row
    with pytest.raises(TypeError):
        rq.ops.dense_gemv_topk(torch.zeros(8, 64, device=dev), torch.zeros(1, 64, device=dev), 1, 0)
    with pytest.raises(ValueError):
        rq.ops.dense_gemv_topk(torch.zeros(8, 64, device=dev, dtype=torch.bfloat16),
                               torch.zeros(9, 64, device=dev, dtype=torch.bfloat16), 1, 0)  # batch > 8
    with pytest.raises(NotImplementedError):
        rq.ops.topk_rows(torch.zeros(2, 10), 1)                              # CPU tensor
    assert rq.DenseIndex(persist_directory=None).search("q") == [] and rq.BM25Index().search("q") == []
    before = rq.ops.launch_count()
    rq.ops.topk_rows(torch.randn(2, 100, device=dev), 5)
    assert rq.ops.launch_count() > before
