"""The oracle against its pins: golden vectors, known answers, the live reference (when mounted)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import bm25_okapi, dense_fusion, philox, router as router_oracle

REFERENCE = Path("/root/reference")


# ------------------------------------------------------------------------------------ BM25
@pytest.fixture(scope="module")
def known(golden_dir):
    with open(golden_dir / "bm25_known_answers.json") as fh:
        return json.load(fh)


def _csr_from_text(corpus):
    docs = [bm25_okapi.tokenize(t) for t in corpus]
    vocab = {}
    ids = [[vocab.setdefault(w, len(vocab)) for w in d] for d in docs]
    off = np.concatenate([[0], np.cumsum([len(d) for d in ids])])
    return vocab, bm25_okapi.OkapiCsr(off, np.concatenate(ids), len(vocab)), docs


def test_bm25_known_answers_literal(known):
    model = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in known["corpus"]])
    assert model.avgdl == pytest.approx(known["avgdl"], rel=1e-15)
    assert model.average_idf == pytest.approx(known["average_idf"], rel=1e-13)
    for word, val in known["idf"].items():
        assert model.idf[word] == pytest.approx(val, rel=1e-12, abs=1e-15)
    for query, want in known["queries"].items():
        got = model.get_scores(bm25_okapi.tokenize(query))
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15, err_msg=query)


def test_bm25_known_answers_csr(known):
    vocab, model, _ = _csr_from_text(known["corpus"])
    for query, want in known["queries"].items():
        terms = [vocab.get(w, -1) for w in bm25_okapi.tokenize(query)]
        np.testing.assert_allclose(model.get_scores(terms), want, rtol=1e-12, atol=1e-15, err_msg=query)


def test_bm25_survey_special_cases(known):
    """idf == 0 stays 0 (df == N/2), negative idf gets the epsilon floor, OOV query -> no result."""
    model = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in known["corpus"]])
    assert model.idf["sun"] == 0.0 and model.idf["bright"] == 0.0
    assert model.idf["the"] == pytest.approx(0.25 * model.average_idf)
    assert model.idf["is"] == pytest.approx(0.25 * model.average_idf)
    assert bm25_okapi.index_search(model.get_scores(["zzz"]), 10) == []
    # duplicates count once per occurrence
    one = model.get_scores(["python", "language"])
    two = model.get_scores(["python", "python", "language"])
    assert two[4] - one[4] == pytest.approx(model.get_scores(["python"])[4])


def test_bm25_literal_equals_csr_random():
    rng = np.random.default_rng(5)
    vocab_n = 300
    docs = [list(rng.zipf(1.3, size=rng.integers(5, 60)) % vocab_n) for _ in range(400)]
    lit = bm25_okapi.OkapiLiteral(docs)
    off = np.concatenate([[0], np.cumsum([len(d) for d in docs])])
    csr = bm25_okapi.OkapiCsr(off, np.concatenate(docs), vocab_n)
    assert csr.average_idf == pytest.approx(lit.average_idf, rel=1e-12)
    for _ in range(20):
        q = list(rng.integers(0, vocab_n + 5, size=rng.integers(1, 9)))
        np.testing.assert_allclose(csr.get_scores(q), lit.get_scores(q), rtol=1e-11, atol=1e-14)


def test_index_search_order_and_positivity():
    scores = np.array([0.0, 2.0, 2.0, -1.0, 3.0, 0.0])
    assert bm25_okapi.index_search(scores, 10) == [(4, 3.0), (1, 2.0), (2, 2.0)]
    assert bm25_okapi.index_search(scores, 2) == [(4, 3.0), (1, 2.0)]


# ---------------------------------------------------------------------------------- fusion
def test_hybrid_search_restatement():
    bm = [(3, 4.0), (7, 2.0), (9, 1.0)]
    de = [(7, 0.9), (1, 0.6), (3, 0.3)]
    rows = dense_fusion.hybrid_search(bm, de, top_k=10)
    want = {3: (4 / 4 + 0.3 / 0.9) / 2, 7: (2 / 4 + 0.9 / 0.9) / 2, 9: (1 / 4 + 0) / 2, 1: (0 + 0.6 / 0.9) / 2}
    assert [r[0] for r in rows] == sorted(want, key=lambda i: (-want[i], i))
    for r in rows:
        assert r[3] == pytest.approx(want[r[0]])
    assert dense_fusion.hybrid_search([], [], 5) == []
    # "max(...) or 1": an all-zero side divides by 1
    rows = dense_fusion.hybrid_search([], [(5, 0.0), (6, 0.0)], 5)
    assert [r[3] for r in rows] == [0.0, 0.0] and [r[0] for r in rows] == [5, 6]
    b, d, ids = dense_fusion.scores_for_router(bm, de, num_passages=6)
    assert len(b) == len(d) == len(ids) == 6 and ids[-2:] == [-1, -1] and b[-1] == 0.0


def test_topk_desc_ties():
    s = np.array([[1.0, 5.0, 5.0, 0.0, 5.0]])
    assert [i for i, _ in dense_fusion.topk_desc(s, 3)[0]] == [1, 2, 4]
    assert [i for i, _ in dense_fusion.topk_desc(s, 10, positive_only=True)[0]] == [1, 2, 4, 0]


# ------------------------------------------------- fusion + BM25Index.search vs the LIVE reference's outputs
@pytest.fixture(scope="module")
def retrieval_gold(golden_dir):
    with open(golden_dir / "retrieval_golden.json") as fh:
        return json.load(fh)


def tie_groups(rows, score_of):
    """[[ids of equal score...], ...] in rank order: the reference's order INSIDE a group is set-iteration /
    introsort dependent, the groups themselves and their order are not."""
    groups = []
    for r in rows:
        if groups and score_of(r) == groups[-1][0]:
            groups[-1][1].add(r[0])
        else:
            groups.append((score_of(r), {r[0]}))
    return [g for _, g in groups]


def _number_pools(case, number):
    pool = case["retrieval_pool_size"]
    return ([(number(i), s) for i, s in case["bm25_pool"][:pool]], [(number(i), s) for i, s in case["dense_pool"][:pool]])


def test_fusion_oracle_matches_live_reference_golden(retrieval_gold):
    """oracle.dense_fusion.hybrid_search / scores_for_router == HybridRetriever.hybrid_search /
    get_scores_for_router of /root/reference (streaming_index.py:464-557) on every stored case, bit for bit."""
    hy = retrieval_gold["hybrid"]
    known = {d: i for i, d in enumerate(hy["document_ids"])}
    ghosts = {}

    def number(doc_id):
        return known[doc_id] if doc_id in known else ghosts.setdefault(doc_id, 10_000 + len(ghosts))

    assert len(hy["cases"]) >= 20
    for case in hy["cases"]:
        bpool, dpool = _number_pools(case, number)
        got = dense_fusion.hybrid_search(bpool, dpool, case["top_k"], known_ids=set(known.values()))
        want = [(number(r[0]), r[1], r[2], r[3]) for r in case["hybrid_search"]]
        assert len(got) == len(want), case["name"]
        assert tie_groups(got, lambda r: r[3]) == tie_groups(want, lambda r: r[3]), case["name"]
        by_id = {r[0]: r for r in want}
        for r in got:                                   # same Python-float arithmetic: exact equality
            assert r == by_id[r[0]], (case["name"], r, by_id[r[0]])
        # get_scores_for_router always pools 50 (:537) and pads with 0.0 / ""
        n = case["num_passages"]
        b50 = [(number(i), s) for i, s in case["bm25_pool"][:50]]
        d50 = [(number(i), s) for i, s in case["dense_pool"][:50]]
        ob, od, oi = dense_fusion.scores_for_router(b50, d50, n, known_ids=set(known.values()))
        ref = case["scores_for_router"]
        assert len(ob) == len(od) == len(oi) == n == len(ref["bm25"]) == len(ref["ids"]) == len(ref["texts"])
        ref_rows = [(number(i) if i != "" else -1, b, d) for i, b, d in zip(ref["ids"], ref["bm25"], ref["dense"])]
        assert sorted(zip(oi, ob, od)) == sorted(ref_rows), case["name"]
        n_real = sum(1 for i in ref["ids"] if i != "")
        assert oi[n_real:] == [-1] * (n - n_real) and ob[n_real:] == [0.0] * (n - n_real) and od[n_real:] == [0.0] * (n - n_real)
        assert all(t == (f"text of {i}" if i else "") for i, t in zip(ref["ids"], ref["texts"]))


def test_index_search_oracle_matches_live_reference_golden(retrieval_gold):
    """oracle.bm25_okapi.tokenize + index_search == BM25Index.search of /root/reference (:150-179): argsort cut,
    the > 0 filter, top_k > N, row -> doc id, duplicate documents skipped on add."""
    n_cases = 0
    for corpus in retrieval_gold["bm25_index_search"]:
        docs = [bm25_okapi.tokenize(t) for t in corpus["texts"]]
        lit = bm25_okapi.OkapiLiteral(docs, k1=corpus["k1"], b=corpus["b"])
        vocab = {}
        ids = [[vocab.setdefault(w, len(vocab)) for w in d] for d in docs]
        off = np.concatenate([[0], np.cumsum([len(d) for d in ids])])
        csr = bm25_okapi.OkapiCsr(off, np.concatenate(ids), len(vocab), k1=corpus["k1"], b=corpus["b"])
        for case in corpus["cases"]:
            toks = bm25_okapi.tokenize(case["query"])
            got = bm25_okapi.index_search(lit.get_scores(toks), case["top_k"])
            want = [(corpus["doc_ids"].index(d), s) for d, s in case["result"]]
            assert len(got) == len(want), (corpus["name"], case["query"])
            assert tie_groups(got, lambda r: r[1]) == tie_groups(want, lambda r: r[1]), (corpus["name"], case["query"])
            assert dict(got) == dict(want)              # identical float64 scores
            got_csr = bm25_okapi.index_search(csr.get_scores([vocab.get(w, -1) for w in toks]), case["top_k"])
            assert [i for i, _ in got_csr] == [i for i, _ in got] or tie_groups(got_csr, lambda r: round(r[1], 9)) == \
                tie_groups(got, lambda r: round(r[1], 9))
            np.testing.assert_allclose([s for _, s in got_csr], [s for _, s in got], rtol=1e-12)
            n_cases += 1
    assert n_cases >= 20


@pytest.mark.skipif(not REFERENCE.exists(), reason="live reference only exists in the build container")
def test_fusion_and_index_search_oracle_match_live_reference_random():
    """300 random pool pairs and 40 random corpora through the live classes (not only the stored cases)."""
    sys.path.insert(0, str(REFERENCE))
    try:
        import rag_uq.streaming_index as ref
    finally:
        sys.path.remove(str(REFERENCE))
    rng = np.random.default_rng(77)
    names = [f"doc{i}" for i in range(150)]
    for trial in range(300):
        r = ref.HybridRetriever()
        held = set(rng.choice(150, size=int(rng.integers(100, 151)), replace=False).tolist())
        for i in held:
            r.documents[names[i]] = ref.Document(id=names[i], text=f"t{i}")
        nb, nd = int(rng.integers(0, 61)), int(rng.integers(0, 61))
        lo_d = float(rng.choice([-1.0, -0.2, 0.0, 0.3]))
        bp = [(int(i), float(s)) for i, s in zip(rng.choice(150, nb, replace=False), np.sort(rng.uniform(0.01, 25, nb))[::-1])]
        dp = [(int(i), float(s)) for i, s in zip(rng.choice(150, nd, replace=False), np.sort(rng.uniform(lo_d, lo_d + 1, nd))[::-1])]
        if trial % 9 == 0:
            dp = [(i, 0.0) for i, _ in dp]
        pool, top_k = int(rng.integers(1, 70)), int(rng.integers(1, 40))
        r.bm25_search = lambda q, k=20, bp=bp: [(names[i], s) for i, s in bp[:k]]
        r.dense_search = lambda q, k=20, dp=dp: [(names[i], s) for i, s in dp[:k]]
        everything = r.hybrid_search("q", 10 ** 6, pool)
        hs = [x.hybrid_score for x in everything]
        if top_k < len(hs) and hs[top_k - 1] == hs[top_k]:
            continue                                     # cut inside a tie group: membership is set-order dependent
        live = [(names.index(x.doc_id), x.bm25_score, x.dense_score, x.hybrid_score) for x in everything[:top_k]]
        got = dense_fusion.hybrid_search(bp[:pool], dp[:pool], top_k, known_ids=held)
        assert tie_groups(got, lambda x: x[3]) == tie_groups(live, lambda x: x[3])
        assert sorted(got) == sorted(live)
    saved = getattr(ref, "BM25Okapi", None)
    ref.BM25Okapi = bm25_okapi.OkapiLiteral      # the un-vendored third-party class (see make_retrieval_golden.py)
    try:
        words = [f"w{i}" for i in range(40)]
        for trial in range(40):
            n = int(rng.integers(1, 60))
            texts = [" ".join(rng.choice(words, size=int(rng.integers(1, 25))).tolist()) for _ in range(n)]
            index = ref.BM25Index()
            index.add_documents([ref.Document(id=f"d{i}", text=t) for i, t in enumerate(texts)])
            lit = bm25_okapi.OkapiLiteral([bm25_okapi.tokenize(t) for t in texts])
            for _ in range(5):
                q = " ".join(rng.choice(words + ["oov"], size=int(rng.integers(1, 6))).tolist())
                k = int(rng.integers(1, 2 * n + 2))
                scores = lit.get_scores(bm25_okapi.tokenize(q))
                ranked = np.sort(scores)[::-1]
                if k < n and ranked[k - 1] == ranked[k] and ranked[k - 1] > 0:
                    continue
                live = [(int(d[1:]), s) for d, s in index.search(q, k)]
                got = bm25_okapi.index_search(scores, k)
                assert tie_groups(got, lambda x: x[1]) == tie_groups(live, lambda x: x[1])
                assert dict(got) == dict(live)
    finally:
        if saved is None:
            del ref.BM25Okapi
        else:
            ref.BM25Okapi = saved


# ---------------------------------------------------------------------------------- router
@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "router_golden.npz")


def _state(gold, tag):
    pre = f"{tag}/state/"
    return {k[len(pre):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(pre)}


@pytest.mark.parametrize("tag", ["h64", "h32"])
def test_router_oracle_matches_golden(gold, tag):
    state = _state(gold, tag)
    b, d = torch.from_numpy(gold[f"{tag}/bm25"]), torch.from_numpy(gold[f"{tag}/dense"])
    g0 = router_oracle.gate(b, d, state, stats_initialized=False)
    assert torch.equal(g0, torch.from_numpy(gold[f"{tag}/gate_batchstat"]))
    vals, idx = router_oracle.hybrid_rerank(b, d, state, False, 10)
    torch.testing.assert_close(vals, torch.from_numpy(gold[f"{tag}/rerank_batchstat_vals"]), rtol=0, atol=0)
    assert torch.equal(idx, torch.from_numpy(gold[f"{tag}/rerank_batchstat_idx"]))

    run = dict(state)
    for name, val in zip(["bm25_mean", "bm25_std", "dense_mean", "dense_std"], gold[f"{tag}/running_stats"]):
        run[name] = torch.tensor(float(val))
    g1 = router_oracle.gate(b, d, run, stats_initialized=True)
    assert torch.equal(g1, torch.from_numpy(gold[f"{tag}/gate_running"]))
    vals, idx = router_oracle.hybrid_rerank(b, d, run, True, 10)
    assert torch.equal(idx, torch.from_numpy(gold[f"{tag}/rerank_running_idx"]))
    bb, bd = torch.from_numpy(gold[f"{tag}/big_bm25"]), torch.from_numpy(gold[f"{tag}/big_dense"])
    vals, idx = router_oracle.hybrid_rerank(bb, bd, run, True, 500)      # k > P clamps to P
    assert vals.shape == (2, 300)
    torch.testing.assert_close(vals, torch.from_numpy(gold[f"{tag}/big_rerank_vals"]), rtol=0, atol=0)

    masks = torch.from_numpy(gold[f"{tag}/mc_masks"]).float()
    for t in range(masks.shape[0]):
        g = router_oracle.gate(b, d, run, True, keep_mask=masks[t])
        torch.testing.assert_close(g, torch.from_numpy(gold[f"{tag}/mc_gates"][t]), rtol=1e-6, atol=1e-7)


def test_router_oracle_mc_aggregation_matches_confidence_math(gold):
    state = _state(gold, "h64")
    b, d = torch.from_numpy(gold["h64/bm25"]), torch.from_numpy(gold["h64/dense"])
    masks = torch.from_numpy(gold["h64/mc_masks"]).float()
    out = router_oracle.mc_dropout(b, d, state, False, masks)
    gates = torch.stack([router_oracle.gate(b, d, state, False, masks[t]) for t in range(masks.shape[0])]).numpy()
    for q in range(b.shape[0]):
        emb = gates[:, q, :]                                  # the T "embeddings" of query q
        centroid = emb.mean(axis=0)                           # confidence.py:196
        dist = np.linalg.norm(emb - centroid, axis=1)         # :199
        assert float(out["variance"][q]) == pytest.approx(float(dist.std()), rel=1e-5)   # :200
        assert float(out["uncertainty"][q]) == pytest.approx(min(1.0, float(dist.std()) / 2.0), rel=1e-5)
        assert int(out["consensus"][q]) == int(np.argmin(dist))
    np.testing.assert_allclose(out["std_w"].numpy(), gates.std(axis=0), rtol=1e-4, atol=1e-7)


def test_router_single_element_is_nan_like_reference(gold):
    state = _state(gold, "h64")
    g = router_oracle.gate(torch.tensor([[1.0]]), torch.tensor([[0.5]]), state, False)
    assert torch.isnan(g).all()      # torch.std of one element (router.py:135) -> NaN, SURVEY 8(a10)


@pytest.mark.skipif(not REFERENCE.exists(), reason="live reference only exists in the build container")
def test_router_oracle_matches_live_reference():
    sys.path.insert(0, str(REFERENCE))
    try:
        from rag_uq.router import RetrievalRouter, RouterConfig
    finally:
        sys.path.remove(str(REFERENCE))
    torch.manual_seed(3)
    live = RetrievalRouter(RouterConfig(hidden_dim=64)).eval()
    state = {k: v.detach().clone() for k, v in live.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    b, d = torch.rand(8, 50, generator=g) * 12, torch.rand(8, 50, generator=g) * 2 - 1
    with torch.no_grad():
        assert torch.equal(live(b, d), router_oracle.gate(b, d, state, False))
        lv, li = live.hybrid_rerank(b, d, top_k=7)
        ov, oi = router_oracle.hybrid_rerank(b, d, state, False, 7)
        assert torch.equal(lv, ov) and torch.equal(li, oi)
        live.train()
        torch.manual_seed(99)
        mask = torch.empty(b.numel(), 64).bernoulli_(0.9)
        torch.manual_seed(99)
        lw = live(b, d, update_stats=False)
        torch.testing.assert_close(lw, router_oracle.gate(b, d, state, False, keep_mask=mask), rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------- philox
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in cases:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_philox_uniform_and_geometry():
    u = philox.uniform_from_bits(np.array([0, 0xFFFFFFFF], dtype=np.uint32))
    assert u[0] == np.float32(2.0 ** -33) and u[1] <= 1.0
    grid, threads, inc = philox.torch_dropout_geometry(6400, 148)
    assert (grid, threads, inc) == (25, 6400, 4)
    grid, threads, inc = philox.torch_dropout_geometry(1024 * 100 * 64, 148)
    assert grid == 1184 and inc == ((6553600 - 1) // (256 * 1184 * 4) + 1) * 4
    keep = philox.keep_mask_torch_layout(6400, seed=123, offset=0, keep_prob=0.9, sm_count=148)
    assert keep.shape == (6400,) and 0.85 < keep.mean() < 0.95


def test_hnsw_oracle_finds_neighbours_on_clustered_data():
    """The HNSW restatement behaves like an HNSW: near-perfect recall on low-dimensional clustered vectors with
    search_ef = 100, lower with ef = 10, deterministic for a fixed seed."""
    from oracle import hnsw
    rng = np.random.default_rng(0)
    centres = rng.normal(size=(12, 24))
    x = (centres[rng.integers(0, 12, 1200)] + 0.3 * rng.normal(size=(1200, 24))).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = x[:40] + 0.05 * rng.normal(size=(40, 24)).astype(np.float32)
    exact = np.argsort(-((q / np.linalg.norm(q, axis=1, keepdims=True)) @ x.T), axis=1)[:, :10]
    rep = hnsw.recall_report(x, q, exact, 10)
    assert rep["recall@10_search_ef_100"] >= 0.98
    assert rep["recall@10_search_ef_10"] <= rep["recall@10_search_ef_100"]
    a, b = hnsw.HnswCosine(24), hnsw.HnswCosine(24)
    a.add(x[:300]); b.add(x[:300])
    assert a.links == b.links
    ids, sims = a.search(x[5], 5)
    assert ids[0] == 5 and sims[0] == pytest.approx(1.0, abs=1e-5) and list(sims) == sorted(sims, reverse=True)
