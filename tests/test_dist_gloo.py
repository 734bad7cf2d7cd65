"""World-size-2 checks of the N>1 host logic on CPU (gloo): row sharding, the global BM25
statistics all-reduce and the candidate all-gather layout the merge kernel consumes."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rag_uq_b200 as rq
        from oracle import dense_fusion
        from rag_uq_b200 import synth

        n, k, n_q = 1001, 5, 3
        lo, hi = rq.shard_rows(n, world, rank)
        # --- global statistics: every rank ends up with the corpus-wide df / N / length
        cdf = synth.zipf_cdf(synth.vocab_size(n), "cpu")
        off, tok = synth.doc_tokens(lo, hi, cdf)
        vocab = cdf.shape[0]
        owner = torch.repeat_interleave(torch.arange(hi - lo), off[1:] - off[:-1])
        pairs = torch.unique(tok.long() * (hi - lo) + owner)
        df_local = torch.bincount(pairs // (hi - lo), minlength=vocab).to(torch.int32)
        df, n_all, len_all = rq.global_bm25_statistics(df_local, hi - lo, int(off[-1]))
        off_all, tok_all = synth.doc_tokens(0, n, cdf)
        owner_all = torch.repeat_interleave(torch.arange(n), off_all[1:] - off_all[:-1])
        df_want = torch.bincount(torch.unique(tok_all.long() * n + owner_all) // n, minlength=vocab).to(torch.int32)
        assert n_all == n and len_all == int(off_all[-1]) and torch.equal(df, df_want)

        # --- candidate exchange: local top-k of this shard's rows, gathered rank-major per query
        g = torch.Generator().manual_seed(5)
        scores = torch.randn(n_q, n, generator=g)                    # same "global" matrix on every rank
        local = dense_fusion.topk_desc(scores[:, lo:hi].double().numpy(), k)
        ls = torch.tensor([[c[1] for c in row] for row in local], dtype=torch.float32)
        li = torch.tensor([[c[0] + lo for c in row] for row in local], dtype=torch.int32)
        gs, gi = rq.gather_candidates(ls, li)
        assert gs.shape == (n_q, world, k) and gi.shape == (n_q, world, k)
        assert torch.equal(gs[:, rank], ls) and torch.equal(gi[:, rank], li)
        want = dense_fusion.topk_desc(scores.double().numpy(), k)
        for q in range(n_q):
            parts = [[(int(gi[q, r, j]), float(gs[q, r, j])) for j in range(k)] for r in range(world)]
            assert dense_fusion.merge_topk(parts, k) == [(i, pytest.approx(s)) for i, s in want[q]]
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_exchange(lib_built):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}
