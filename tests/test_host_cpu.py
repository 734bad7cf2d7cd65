"""Host-side checks that need no GPU: the C ABI exports what include/ragb200.h declares, the
ops refuse to run anywhere but on CUDA, generators are deterministic, host logic of the
drop-in classes behaves like the reference's."""
import ctypes
import json
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def rq(lib_built):
    import rag_uq_b200
    return rag_uq_b200


def declared_functions():
    text = (ROOT / "include" / "ragb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ragb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(rq):
    names = declared_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(str(rq._lib.LIB_PATH))
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ragb200.h but not exported"
    assert set(names) == set(rq._lib.SIGNATURES), "ctypes signature table and header disagree"
    assert lib.ragb_abi_version() == 1


def test_library_is_sm100a_with_blackwell_instructions(rq):
    """The shipped code object holds tcgen05 / TMA SASS and nothing for other architectures."""
    out = subprocess.run(["cuobjdump", "-lelf", str(rq._lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", str(rq._lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback(rq):
    with pytest.raises(NotImplementedError):
        rq.ops.topk_rows(torch.zeros(2, 8), 2)
    with pytest.raises(NotImplementedError):
        rq.ops.dense_gemv_topk(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(1, 64, dtype=torch.bfloat16), 1, 0)
    router = rq.RetrievalRouter()
    with pytest.raises(NotImplementedError):
        router(torch.randn(2, 5), torch.randn(2, 5))
    with pytest.raises(NotImplementedError):
        rq.HybridEngine(None, torch.zeros(4, 64, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        # the C ABI itself refuses when there is no sm_100 device
        rc = rq._lib.lib.ragb_topk_merge(None, None, 1, 1, 1, 1, None, None, None)
        assert rc in (rq._lib.RAGB_ECUDA, rq._lib.RAGB_EARCH)
        assert rq._lib.last_error()
        index = rq.BM25Index()
        index.add_documents([rq.Document(id="a", text="x y z")])
        with pytest.raises(RuntimeError):
            index.search("x")


def test_product_package_does_not_import_oracle():
    pkg = ROOT / "efficient-rag-with-learned-retrieval-and-uncertainty-quantification_b200"
    for path in (list(pkg.glob("*.py")) + list(pkg.glob("csrc/*")) + [ROOT / "rag_uq_b200" / "__init__.py"]
                 + list((ROOT / "scripts").glob("*.py"))):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path


def test_router_state_dict_is_reference_compatible(rq, golden_dir):
    gold = np.load(golden_dir / "router_golden.npz")
    state = {k[len("h64/state/"):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("h64/state/")}
    router = rq.RetrievalRouter()
    assert set(router.state_dict()) == set(state)
    router.load_state_dict(state)                       # strict: same keys, same shapes
    assert router.stats_initialized is False            # plain attribute, not in the checkpoint
    assert router._count_params() == 321
    assert rq.RetrievalRouter(rq.RouterConfig(hidden_dim=32))._count_params() == 161
    with pytest.raises(ValueError):
        rq.RetrievalRouter(rq.RouterConfig(num_layers=3))
    with pytest.raises(ValueError):
        rq.RetrievalRouter(rq.RouterConfig(use_batch_norm=True))


def test_reference_checkpoint_unpickles_through_the_module_alias(rq, golden_dir):
    """tests/golden/router_checkpoint.pt was written by the LIVE reference (RouterTrainer.save_checkpoint): its pickled
    rag_uq.router.RouterConfig resolves to rag_uq_b200's class through the documented alias and the state dict loads."""
    import sys
    import rag_uq_b200.router as router_module
    saved = {k: sys.modules.get(k) for k in ("rag_uq", "rag_uq.router")}
    sys.modules.setdefault("rag_uq", type(sys)("rag_uq"))
    sys.modules["rag_uq.router"] = router_module
    try:
        ckpt = torch.load(golden_dir / "router_checkpoint.pt", map_location="cpu", weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert type(ckpt["config"]) is rq.RouterConfig and ckpt["config"].hidden_dim == 32
    router = rq.RetrievalRouter(ckpt["config"])
    missing, unexpected = router.load_state_dict(ckpt["model_state_dict"], strict=True)
    assert not missing and not unexpected and len(ckpt["train_losses"]) == 6
    assert float(router.bm25_std) != 1.0                      # the EMA statistics of the training steps came along


def test_synth_is_deterministic_and_shardable(rq):
    from rag_uq_b200 import synth
    a = synth.passage_embeddings(100, 228, 64, "cpu")
    b = synth.passage_embeddings(0, 300, 64, "cpu")[100:228]
    assert torch.equal(a, b)
    cdf = synth.zipf_cdf(synth.vocab_size(1000), "cpu")
    off1, tok1 = synth.doc_tokens(10, 60, cdf)
    off2, tok2 = synth.doc_tokens(0, 60, cdf)
    assert torch.equal(tok1, tok2[int(off2[10]):])
    lens = (off2[1:] - off2[:-1])
    assert 20 <= int(lens.min()) and int(lens.max()) <= 300
    q1 = synth.make_queries(5, 1000, 64, cdf, "cpu")
    q2 = synth.make_queries(5, 1000, 64, cdf, "cpu")
    assert torch.equal(q1.q_terms, q2.q_terms) and torch.equal(q1.q_emb, q2.q_emb)
    assert q1.q_off.tolist() == [0, 8, 16, 24, 32, 40]
    # the query's tokens really come from its source passage
    off, tok = synth.doc_tokens(0, 1000, cdf)
    for i in range(5):
        src = int(q1.source_rows[i])
        doc = set(tok[int(off[src]):int(off[src + 1])].tolist())
        assert all(t in doc or t >= cdf.shape[0] for t in q1.q_terms[8 * i:8 * i + 8].tolist())


def test_large_check_streams_the_same_oracle(rq):
    """oracle.large_check (what `bench.py --verify` and the 1M GPU test use on corpora too large for the CSR oracle)
    == OkapiCsr + exact dense + oracle fusion + oracle rerank on a corpus small enough for both."""
    from oracle import bm25_okapi, dense_fusion, large_check, router as router_oracle
    from rag_uq_b200 import synth
    n, dim, n_q, pool, k = 3000, 64, 6, 50, 10
    vocab = synth.vocab_size(n)
    cdf = synth.zipf_cdf(vocab, "cpu")
    qb = synth.make_queries(n_q, n, dim, cdf, "cpu")
    terms = qb.q_terms.view(n_q, -1).tolist()
    terms[2] = terms[2][:3] + terms[2][:3]            # duplicates count per occurrence
    torch.manual_seed(7)
    lin1, lin2 = torch.nn.Linear(3, 64), torch.nn.Linear(64, 1)
    state = {"scorer.0.weight": lin1.weight.detach(), "scorer.0.bias": lin1.bias.detach(), "scorer.3.weight": lin2.weight.detach(),
             "scorer.3.bias": lin2.bias.detach(), "bm25_mean": torch.tensor(8.0), "bm25_std": torch.tensor(6.0),
             "dense_mean": torch.tensor(0.2), "dense_std": torch.tensor(0.3)}
    recs, info = large_check.run_synthetic_check(synth, "cpu", n, dim, terms, qb.q_emb, pool, k, state, True, chunk_docs=700)
    off, tok = synth.doc_tokens(0, n, cdf)
    okapi = bm25_okapi.OkapiCsr(off.numpy(), tok.numpy(), vocab)
    assert info["average_idf"] == pytest.approx(okapi.average_idf, rel=1e-13) and info["avgdl"] == okapi.avgdl
    emb = synth.passage_embeddings(0, n, dim, "cpu").float().numpy()
    dense = dense_fusion.dense_scores(emb, qb.q_emb.float().numpy())
    for q in range(n_q):
        bm = bm25_okapi.index_search(okapi.get_scores(terms[q]), pool)
        de = dense_fusion.topk_desc(dense[q:q + 1], pool)[0]
        assert [i for i, _ in recs[q]["bm25_pool"]] == [i for i, _ in bm]
        np.testing.assert_allclose([s for _, s in recs[q]["bm25_pool"]], [s for _, s in bm], rtol=1e-13)
        assert [i for i, _ in recs[q]["dense_pool"]] == [i for i, _ in de]
        want = dense_fusion.hybrid_search(bm, de, k)
        assert [r[0] for r in recs[q]["fused"]] == [r[0] for r in want]
        sb, sd, ids = dense_fusion.scores_for_router(bm, de, k)
        vals, order = router_oracle.hybrid_rerank(torch.tensor([sb]), torch.tensor([sd]), state, True, k)
        assert recs[q]["rerank_ids"] == [ids[j] for j in order[0].tolist()]
    # top_desc: ties in index order, positivity filter, gap report
    idx, val, gap = large_check.top_desc(np.array([0.0, 2.0, 2.0, -1.0, 3.0, 0.0]), 3)
    assert idx.tolist() == [4, 1, 2] and gap == 1.0
    idx, _, _ = large_check.top_desc(np.array([0.0, 2.0, 2.0, -1.0, 3.0, 0.0]), 5, positive_only=True)
    assert idx.tolist() == [4, 1, 2]
    exact, explained = large_check.compare_ranking([4, 2, 1], [3.0, 2.0, 2.0], [4, 1, 2], [3.0, 2.0, 2.0], {4: 3.0, 1: 2.0, 2: 2.0}, 1e-6, 0)
    assert (exact, explained) == (False, True)
    assert large_check.compare_ranking([4, 0], [3.0, 2.0], [4, 1], [3.0, 2.0], {4: 3.0, 1: 2.0, 0: 0.0}, 1e-6, 0) == (False, False)


def test_shard_rows_partition():
    import rag_uq_b200 as rq
    for n, world in [(10, 3), (10_000_000, 8), (7, 8), (1, 1)]:
        parts = [rq.shard_rows(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_seed_slices_cover_every_query_once():
    """HybridEngine.local_pools at world > 1: every shard seeds one contiguous slice of the batch, the MAX all-reduce
    distributes the bounds - so the slices must partition [0, n) for any batch size and world size."""
    from rag_uq_b200.engine import seed_slice
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8, 16):
            slices = [seed_slice(n, r, world) for r in range(world)]
            covered = [q for q0, q1 in slices for q in range(q0, q1)]
            assert covered == list(range(n)), (n, world, slices)
            assert all(0 <= q0 <= q1 <= n for q0, q1 in slices)


def test_bound_table_cache_keeps_several_grids(rq):
    """router.full_fusion_table / full_fusion_envelope: the fallback of the threshold search alternates between a few
    (b_cap, d_hi) grids; each stays cached until a weight or statistic changes (a one-entry cache recomputed the host
    table - tens of milliseconds - on every alternation)."""
    torch.manual_seed(3)
    router = rq.RetrievalRouter()
    router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.2); router.dense_std.fill_(0.3)
    router.stats_initialized = True
    a = router.full_fusion_table(32.0, 1.0)
    b = router.full_fusion_table(64.0, 1.015625)
    assert router.full_fusion_table(32.0, 1.0) is a and router.full_fusion_table(64.0, 1.015625) is b
    e = router.full_fusion_envelope(32.0, 1.0, 64, 32)
    assert router.full_fusion_envelope(32.0, 1.0, 64, 32) is e
    router.bm25_mean.fill_(9.0)                                   # a statistic moved: every cached grid is stale
    c = router.full_fusion_table(32.0, 1.0)
    assert c is not a and len(router.__dict__["_ff_cache"]) == 1
    assert not torch.equal(c, a)


def test_streaming_index_checkpoint_and_resume(rq, tmp_path):
    """StreamingIndex keeps the reference's checkpoint semantics (streaming_index.py:593-679)."""

    class FakeRetriever:
        def __init__(self):
            self.batches, self.documents = [], {}

        def add_documents(self, docs):
            self.batches.append([d.id for d in docs])
            self.documents.update({d.id: d for d in docs})
            return {}

        def __len__(self):
            return len(self.documents)

    corpus = tmp_path / "corpus.jsonl"
    lines = [json.dumps({"id": f"d{i}", "text": f"text {i}", "title": f"t{i}"}) for i in range(7)]
    lines.insert(3, "{not json")                         # skipped with a warning, still counts as an offset
    lines.insert(5, json.dumps({"id": "no-text"}))       # KeyError -> skipped
    corpus.write_text("\n".join(lines) + "\n")
    ckpt = tmp_path / "state" / "ckpt.json"

    fake = FakeRetriever()
    index = rq.StreamingIndex(fake, checkpoint_path=str(ckpt), batch_size=3)
    assert index.get_progress() == {"last_offset": 0, "total_indexed": 0, "files_completed": [], "retriever_size": 0}
    gen = index.stream_from_jsonl(str(corpus))
    assert next(gen) == 3
    saved = json.loads(ckpt.read_text())
    assert saved["last_offset"] == 3 and saved["total_indexed"] == 3      # committed right after the third line
    # crash here; a new process resumes from the checkpoint and does not re-add the first batch
    fake2 = FakeRetriever()
    index2 = rq.StreamingIndex(fake2, checkpoint_path=str(ckpt), batch_size=3)
    assert list(index2.stream_from_jsonl(str(corpus))) == [3, 1]
    assert fake2.batches == [["d3", "d4", "d5"], ["d6"]]
    final = json.loads(ckpt.read_text())
    assert final == {"last_offset": 9, "total_indexed": 7, "files_completed": [str(corpus)]}
    with pytest.raises(FileNotFoundError):
        list(index2.stream_from_jsonl(str(tmp_path / "missing.jsonl")))
    # resume=False starts over
    fake3 = FakeRetriever()
    index3 = rq.StreamingIndex(fake3, checkpoint_path=str(tmp_path / "other.json"), batch_size=100)
    assert list(index3.stream_from_jsonl(str(corpus), resume=False)) == [7]


def test_bm25_index_host_state_and_pickle_schema(rq, tmp_path):
    path = tmp_path / "idx" / "bm25.pkl"
    index = rq.BM25Index(persist_path=str(path), k1=1.2, b=0.5)
    docs = [rq.Document(id="a", text="The Sky  is BLUE"), rq.Document(id="b", text="the sun", title="Sun", metadata={"x": 1})]
    assert index.add_documents(docs) == 2
    assert index.add_documents(docs) == 0                           # ids already present (:134)
    assert index.tokenized_corpus[0] == ["the", "sky", "is", "blue"] and len(index) == 2
    assert index.get_document("b").title == "Sun" and index.get_document("zz") is None
    import pickle
    payload = pickle.loads(path.read_bytes())
    assert set(payload) == {"documents", "doc_ids", "tokenized_corpus", "k1", "b"}
    assert payload["documents"]["b"] == {"id": "b", "text": "the sun", "title": "Sun", "metadata": {"x": 1}}
    again = rq.BM25Index(persist_path=str(path))
    assert again.doc_ids == ["a", "b"] and again.k1 == 1.2 and again.b == 0.5 and again.vocab == index.vocab
    q_rows = [[again.vocab.get(t, -1) for t in again._tokenize(q)] for q in ["the moon", ""]]
    assert q_rows == [[again.vocab["the"], -1], []]
    assert rq.BM25Index().search("anything") == []                   # empty index -> [] (:165-166)
    d = rq.Document.from_dict({"id": "q", "text": "t"})
    assert d.title is None and d.to_dict() == {"id": "q", "text": "t", "title": "", "metadata": {}}


def test_confidence_host_math(rq):
    from rag_uq_b200.confidence import embedding_variance, lexical_diversity
    emb = np.array([[0.0, 0.0], [2.0, 0.0], [0.0, 2.0], [2.0, 2.0]])
    var, centroid, dist = embedding_variance(emb)
    assert np.allclose(centroid, [1, 1]) and np.allclose(dist, np.sqrt(2)) and var == pytest.approx(0.0)
    assert lexical_diversity(["a b", "a c"]) == 0.75 and lexical_diversity([]) == 1.0

    class LLM:
        def __init__(self):
            self.n = 0

        def generate(self, **kw):
            self.n += 1
            return {"response": ["Paris", "Paris", "London", ""][self.n % 4]}

    class Enc:
        def encode(self, texts):
            return np.array([[1.0, 0.0] if t == "Paris" else [0.0, 1.0] for t in texts])

    res = rq.MCDropoutConfidence(LLM(), n_samples=8, encoder=Enc()).get_confidence_interval("p", "c", "q")
    assert res.consensus_answer == "Paris" and 0.0 <= res.confidence <= 1.0
    assert res.uncertainty_score == pytest.approx(min(1.0, res.embedding_variance / 2))
    assert res.metadata["n_samples"] == 6


@pytest.mark.parametrize("hidden,scale", [(64, 1.0), (32, 4.0), (128, 1.0), (16, 20.0)])
def test_full_fusion_gate_bounds_are_proven_bounds(rq, hidden, scale):
    """router.full_fusion_bounds: lo <= gate <= hi on every cell (checked against the fp32 torch gate on 400k random
    points), last bm25 row (0, 1), and the fused-score bound the kernel derives from it never undercuts the score."""
    from rag_uq_b200.router import full_fusion_bounds
    torch.manual_seed(hidden)
    lin1, lin2 = torch.nn.Linear(3, hidden), torch.nn.Linear(hidden, 1)
    w1, b1 = (lin1.weight.detach() * scale).numpy(), (lin1.bias.detach() * scale).numpy()
    w2, b2 = (lin2.weight.detach() * scale).reshape(-1).numpy(), lin2.bias.detach().numpy()
    stats = np.array([8.0, 6.0, 0.2, 0.3], dtype=np.float32)
    b_cap, d_hi, n_b, n_d = 64.0, 1.02, 128, 64
    table = full_fusion_bounds(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d).view(np.uint32)
    assert table.shape == (n_b, n_d)
    lo = (table << 16).view(np.float32)
    hi = (table & np.uint32(0xFFFF0000)).view(np.float32)
    assert (lo[-1] == 0).all() and (hi[-1] == 1).all() and (lo <= hi).all() and (lo >= 0).all() and (hi <= 1).all()
    rng = np.random.default_rng(3)
    n = 400_000
    b = rng.uniform(-2.0, b_cap * 1.1, n).astype(np.float32)
    d = rng.uniform(-d_hi, d_hi, n).astype(np.float32)
    bt, dt = torch.tensor(b), torch.tensor(d)
    bn = (bt - stats[0]) / (torch.tensor(stats[1]) + 1e-6)
    dn = (dt - stats[2]) / (torch.tensor(stats[3]) + 1e-6)
    feats = torch.stack([bn, dn, dn - bn], -1)
    gate = torch.sigmoid(torch.relu(feats @ torch.tensor(w1).T + torch.tensor(b1)) @ torch.tensor(w2) + torch.tensor(b2))
    fused = (gate * dt + (1 - gate) * bt).numpy()
    gate = gate.numpy()
    ib = np.floor(b * np.float32(n_b / b_cap)).astype(np.int64)
    ib = np.where((ib < 0) | (ib > n_b - 1), n_b - 1, ib)           # the kernel's clamp: negative / huge -> last row
    idx = np.clip(np.floor((d + d_hi) * (n_d / (2 * d_hi))).astype(np.int64), 0, n_d - 1)
    assert (lo[ib, idx] <= gate).all() and (gate <= hi[ib, idx]).all()
    bound = b + np.where(d <= b, lo[ib, idx], hi[ib, idx]) * (d - b)
    assert (bound >= fused - 1e-5 * np.abs(fused) - 1e-6).all()
    # a degenerate router (NaN weight) yields the always-valid table
    bad = full_fusion_bounds(w1 * np.nan, b1, w2, b2, stats, b_cap, d_hi, 8, 4).view(np.uint32)
    assert ((bad << 16).view(np.float32) == 0).all() and ((bad & np.uint32(0xFFFF0000)).view(np.float32) == 1).all()


@pytest.mark.parametrize("hidden,scale", [(64, 1.0), (16, 20.0), (32, 4.0)])
def test_full_fusion_envelope_is_a_monotone_proven_bound(rq, hidden, scale):
    """router.full_fusion_envelope: the maximum of E[ib, id_lo..id_hi] bounds the fused score of every (bm25, dense) with
    bm25 at or below row ib and dense inside the column range - the stopping rule of the threshold-algorithm
    full-fusion search - and E is monotone in the BM25 direction."""
    from rag_uq_b200.router import full_fusion_envelope
    torch.manual_seed(hidden)
    lin1, lin2 = torch.nn.Linear(3, hidden), torch.nn.Linear(hidden, 1)
    w1, b1 = (lin1.weight.detach() * scale).numpy(), (lin1.bias.detach() * scale).numpy()
    w2, b2 = (lin2.weight.detach() * scale).reshape(-1).numpy(), lin2.bias.detach().numpy()
    stats = np.array([8.0, 6.0, 0.2, 0.3], dtype=np.float32)
    b_cap, d_hi, n_b, n_d = 32.0, 1.02, 128, 64
    env = full_fusion_envelope(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)
    assert env.shape == (n_b, n_d) and (np.diff(env, axis=0) >= 0).all()
    rng = np.random.default_rng(1)
    n = 200_000
    b = rng.uniform(0, b_cap * 0.999, n).astype(np.float32)
    d = rng.uniform(-d_hi, d_hi, n).astype(np.float32)
    bt, dt = torch.tensor(b), torch.tensor(d)
    bn = (bt - stats[0]) / (torch.tensor(stats[1]) + 1e-6)
    dn = (dt - stats[2]) / (torch.tensor(stats[3]) + 1e-6)
    feats = torch.stack([bn, dn, dn - bn], -1)
    gate = torch.sigmoid(torch.relu(feats @ torch.tensor(w1).T + torch.tensor(b1)) @ torch.tensor(w2) + torch.tensor(b2))
    fused = (gate * dt + (1 - gate) * bt).numpy()
    ib = np.clip(np.floor(b * (n_b / b_cap)).astype(int), 0, n_b - 1)
    idc = np.clip(np.floor((d + d_hi) * (n_d / (2 * d_hi))).astype(int), 0, n_d - 1)
    assert (env[ib, idc] >= fused - 1e-5).all()
    # ... and of everything BELOW the row: a point is also bounded by any row above its own, in its own column, and by
    # the maximum over any column range that contains its column
    up_b = np.minimum(ib + rng.integers(0, 20, n), n_b - 1)
    assert (env[up_b, idc] >= fused - 1e-5).all()
    lo_c, hi_c = np.maximum(idc - rng.integers(0, 5, n), 0), np.minimum(idc + rng.integers(0, 5, n), n_d - 1)
    ranged = np.array([env[r, a:c + 1].max() for r, a, c in zip(up_b[:5000], lo_c[:5000], hi_c[:5000])])
    assert (ranged >= fused[:5000] - 1e-5).all()


def test_shard_format_round_trip_on_cpu(rq, tmp_path):
    """N2: the raw-array shard directory reproduces every array bit for bit (memmap read, no parsing); loading
    without finalize needs no GPU."""
    g = torch.Generator().manual_seed(5)
    vocab, n_docs, nnz, dim = 50, 40, 300, 64
    counts = torch.bincount(torch.randint(0, vocab, (nnz,), generator=g), minlength=vocab)
    term_off = torch.zeros(vocab + 1, dtype=torch.int64)
    term_off[1:] = torch.cumsum(counts, 0)
    shard = rq.SparseShard(term_off, torch.randint(0, n_docs, (nnz,), generator=g, dtype=torch.int32),
                           torch.randint(1, 900, (nnz,), generator=g, dtype=torch.int16),
                           torch.randint(20, 300, (n_docs,), generator=g, dtype=torch.int32), counts.to(torch.int32),
                           n_docs=n_docs, vocab=vocab, id_base=1000, k1=1.2, b=0.6, epsilon=0.3)
    passages = torch.randn(n_docs, dim, generator=g).to(torch.bfloat16)
    d = rq.save_shard(tmp_path / "shard0", shard, passages, id_base=1000)
    meta = json.loads((d / "meta.json").read_text())
    assert meta["format"] == "rag_uq_b200.shard" and meta["sparse"]["nnz"] == nnz and meta["passages"]["dim"] == dim
    assert (d / "passages.bin").stat().st_size == n_docs * dim * 2 and (d / "post_tf.bin").stat().st_size == nnz * 2
    got, emb, base = rq.load_shard(d, "cpu", finalize=False)
    assert base == 1000 and (got.k1, got.b, got.epsilon, got.n_docs, got.vocab) == (1.2, 0.6, 0.3, n_docs, vocab)
    for name in ("term_off", "post_doc", "post_tf", "doc_len", "df"):
        assert torch.equal(getattr(got, name), getattr(shard, name)), name
    assert torch.equal(emb.view(torch.int16), passages.view(torch.int16))
    # a dense-only shard and a foreign directory
    d2 = rq.save_shard(tmp_path / "dense_only", None, passages)
    sp, emb2, _ = rq.load_shard(d2, "cpu", finalize=False)
    assert sp is None and torch.equal(emb2.view(torch.int16), passages.view(torch.int16))
    (tmp_path / "bad").mkdir()
    (tmp_path / "bad" / "meta.json").write_text(json.dumps({"format": "something else"}))
    with pytest.raises(ValueError):
        rq.load_shard(tmp_path / "bad", "cpu")


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` needs no GPU: it times the oracle port of the reference's per-query path on the
    host cores and prints ONE JSON line with the keys the driver reads (same metric / unit / config as our arm)."""
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-docs", "300"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"] == "hybrid top-10 queries/sec over 10Mx768 passages" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "passages" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 stay silent under torchrun
    env = dict(**__import__("os").environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
