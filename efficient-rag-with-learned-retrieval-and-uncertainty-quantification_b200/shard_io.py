"""On-disk format of one row shard ("next" row N2 of SURVEY.md section 8f).

The reference persists its sparse index as ONE pickle of Python lists (documents, ids, tokenised
corpus; rag_uq/streaming_index.py:185-201) that is re-tokenised into ``BM25Okapi`` on every load
(:219-220), and its dense index inside ChromaDB.  ``BM25Index`` here still reads and writes that
pickle so a reference deployment can switch over; THIS module is the format for corpora that do not fit
a pickle: a directory of raw little-endian arrays that ``numpy.memmap`` maps without parsing,

    meta.json      format version, shapes, dtypes, k1 / b / epsilon, id_base, corpus statistics
    term_off.bin   int64  [vocab + 1]     postings of term t are [term_off[t], term_off[t+1])
    post_doc.bin   int32  [nnz]           local rows, ascending per term
    post_tf.bin    uint16 [nnz]
    doc_len.bin    int32  [n_docs]
    df.bin         int32  [vocab]         LOCAL document frequencies (the global ones are re-reduced at load)
    passages.bin   bf16   [n_docs, dim]   unit rows (optional)

exactly the HBM layout of DESIGN.md section 3, so loading is a chunked host-to-device copy and the
derived arrays (idf, norm, dense tf table, impact bounds) are recomputed on the GPU from the global
statistics of whatever set of shards is loaded together (so shards can be re-grouped across ranks).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import numpy as np
import torch

FORMAT = "rag_uq_b200.shard"
VERSION = 1
_FILES = {"term_off": np.int64, "post_doc": np.int32, "post_tf": np.uint16, "doc_len": np.int32, "df": np.int32}
_CHUNK = 1 << 28   # bytes per host-to-device copy


def _write(path: Path, t: torch.Tensor, dtype) -> None:
    """Stream a device tensor to ``path`` in chunks (bf16 / uint16 go through their int16 bit pattern)."""
    flat = t.reshape(-1)
    if flat.dtype in (torch.bfloat16, torch.float16):
        flat = flat.view(torch.int16)
    step = max(1, _CHUNK // max(1, flat.element_size()))
    with open(path, "wb") as fh:
        for lo in range(0, flat.numel(), step):
            fh.write(flat[lo:lo + step].cpu().numpy().astype(np.dtype(dtype).newbyteorder("<"), copy=False).tobytes())


def _read(path: Path, dtype, shape, device, torch_dtype) -> torch.Tensor:
    count = int(np.prod(shape))
    out = torch.empty(count, dtype=torch_dtype, device=device)
    if count == 0:
        return out.reshape(shape)
    mm = np.memmap(path, dtype=np.dtype(dtype).newbyteorder("<"), mode="r", shape=(count,))
    step = max(1, _CHUNK // mm.dtype.itemsize)
    view = out.view(torch.int16) if torch_dtype in (torch.bfloat16, torch.float16) else out
    for lo in range(0, count, step):
        chunk = np.array(mm[lo:lo + step])   # one copy out of the page cache into writable memory
        if chunk.dtype == np.uint16:
            chunk = chunk.view(np.int16)
        view[lo:lo + step].copy_(torch.from_numpy(chunk), non_blocking=False)
    return out.reshape(shape)


def save_shard(directory, sparse=None, passages: Optional[torch.Tensor] = None, id_base: int = 0) -> Path:
    """Write one shard.  ``sparse``: a ``SparseShard`` (single segment) or None; ``passages``: bf16 [rows, dim] or None."""
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    meta = {"format": FORMAT, "version": VERSION, "id_base": int(id_base), "sparse": None, "passages": None}
    if sparse is not None:
        if not hasattr(sparse, "term_off"):
            raise TypeError("save_shard takes a single-segment SparseShard (merge the segments of a SegmentedIndex first)")
        meta["sparse"] = {"n_docs": int(sparse.n_docs), "vocab": int(sparse.vocab), "nnz": int(sparse.nnz),
                          "k1": float(sparse.k1), "b": float(sparse.b), "epsilon": float(sparse.epsilon),
                          "total_len": int(sparse.doc_len.sum())}
        for name, dtype in _FILES.items():
            _write(d / f"{name}.bin", getattr(sparse, name), dtype)
    if passages is not None:
        if passages.dtype != torch.bfloat16 or passages.dim() != 2:
            raise TypeError("passages must be a bf16 [rows, dim] tensor")
        meta["passages"] = {"rows": int(passages.shape[0]), "dim": int(passages.shape[1]), "dtype": "bfloat16"}
        _write(d / "passages.bin", passages.contiguous(), np.uint16)
    (d / "meta.json").write_text(json.dumps(meta, indent=1))
    return d


def load_shard(directory, device, finalize: bool = True, group=None):
    """-> (SparseShard | None, passages | None, id_base).  ``finalize`` computes idf / norm / table rows on the
    device (all-reducing df, N and the total length over ``group`` when torch.distributed is initialised)."""
    from .engine import global_bm25_statistics
    from .sparse import SparseShard

    d = Path(directory)
    meta = json.loads((d / "meta.json").read_text())
    if meta.get("format") != FORMAT or int(meta.get("version", -1)) > VERSION:
        raise ValueError(f"{d}: not a {FORMAT} directory of version <= {VERSION}")
    sparse, passages = None, None
    sm = meta["sparse"]
    if sm is not None:
        shapes = {"term_off": (sm["vocab"] + 1,), "post_doc": (sm["nnz"],), "post_tf": (sm["nnz"],),
                  "doc_len": (sm["n_docs"],), "df": (sm["vocab"],)}
        tdt = {"term_off": torch.int64, "post_doc": torch.int32, "post_tf": torch.int16, "doc_len": torch.int32,
               "df": torch.int32}
        arrays = {name: _read(d / f"{name}.bin", dtype, shapes[name], device, tdt[name]) for name, dtype in _FILES.items()}
        sparse = SparseShard(arrays["term_off"], arrays["post_doc"], arrays["post_tf"], arrays["doc_len"], arrays["df"],
                             n_docs=sm["n_docs"], vocab=sm["vocab"], id_base=meta["id_base"], k1=sm["k1"], b=sm["b"],
                             epsilon=sm["epsilon"])
        if finalize:
            df, n_all, len_all = global_bm25_statistics(sparse.df, sm["n_docs"], sm["total_len"], group)
            sparse.finalize(df, n_all, len_all, group)
    pm = meta["passages"]
    if pm is not None:
        passages = _read(d / "passages.bin", np.uint16, (pm["rows"], pm["dim"]), device, torch.bfloat16)
    return sparse, passages, int(meta["id_base"])


def save_engine(directory, engine) -> Path:
    return save_shard(directory, engine.sparse, engine.passages, engine.id_base)


def load_engine(directory, device, group=None, mma_variant: int = 3):
    from .engine import HybridEngine
    sparse, passages, id_base = load_shard(directory, device, True, group)
    return HybridEngine(sparse, passages, id_base=id_base, group=group, mma_variant=mma_variant)
