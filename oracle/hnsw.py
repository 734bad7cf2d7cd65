"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): an HNSW index restated on the CPU, used to report the
recall@k of the reference's approximate dense path against this repository's exact search.

The reference answers dense queries through ChromaDB (`rag_uq/streaming_index.py:254-263, 355-359`:
``collection.query`` on a collection created with ``metadata={"hnsw:space": "cosine"}``), i.e. through
hnswlib's HNSW graph with ChromaDB's defaults.  Neither chromadb nor hnswlib is installable in this image
(requirements.txt:7 ``chromadb>=0.4.0``; no wheel, no network), so the algorithm is restated here from its
publication - Malkov & Yashunin, "Efficient and robust approximate nearest neighbor search using
Hierarchical Navigable Small World graphs" (Alg. 1 insert, Alg. 2 search-layer, Alg. 4 heuristic
neighbour selection, Alg. 5 k-NN search) - with hnswlib's conventions: M neighbours per node on the
upper layers and 2M on layer 0, level = floor(-ln(U) / ln(M)), heuristic selection without
"extend candidates" / "keep pruned", cosine space = 1 - <a, b> on unit vectors, and ChromaDB's default
parameters M = 16, construction_ef = 100, search_ef = 100 (older ChromaDB releases searched with ef = 10;
``recall_report`` takes both).  **Parity unpinned**: there is no fixture of ChromaDB output to pin it to;
the numbers it yields are "an HNSW with ChromaDB's parameters", not ChromaDB itself.
"""
from __future__ import annotations

import heapq
import math
from typing import List, Sequence, Tuple

import numpy as np


class HnswCosine:
    def __init__(self, dim: int, m: int = 16, ef_construction: int = 100, seed: int = 100):
        self.dim, self.m, self.m0, self.efc = dim, m, 2 * m, max(ef_construction, m)
        self.mult = 1.0 / math.log(m)
        self.rng = np.random.default_rng(seed)
        self.vecs = np.zeros((0, dim), dtype=np.float32)
        self.links: List[List[List[int]]] = []   # links[node][level] = neighbour ids
        self.entry, self.max_level = -1, -1

    # ---- distances ------------------------------------------------------------------------
    def _dist(self, q: np.ndarray, ids: Sequence[int]) -> np.ndarray:
        return 1.0 - self.vecs[np.asarray(ids, dtype=np.int64)] @ q

    # ---- Alg. 2: greedy best-first search of one layer --------------------------------------
    def _search_layer(self, q: np.ndarray, entries: List[Tuple[float, int]], ef: int, level: int) -> List[Tuple[float, int]]:
        visited = {e for _, e in entries}
        cand = list(entries)                      # min-heap by distance
        heapq.heapify(cand)
        best = [(-d, e) for d, e in entries]      # max-heap (negated) of the ef closest so far
        heapq.heapify(best)
        while cand:
            d, c = heapq.heappop(cand)
            if d > -best[0][0] and len(best) >= ef:
                break
            fresh = [n for n in self.links[c][level] if n not in visited]
            if not fresh:
                continue
            visited.update(fresh)
            for dn, n in zip(self._dist(q, fresh).tolist(), fresh):
                if len(best) < ef or dn < -best[0][0]:
                    heapq.heappush(cand, (dn, n))
                    heapq.heappush(best, (-dn, n))
                    if len(best) > ef:
                        heapq.heappop(best)
        return sorted((-d, e) for d, e in best)

    # ---- Alg. 4: heuristic neighbour selection (hnswlib getNeighborsByHeuristic2) ------------
    def _select(self, cands: List[Tuple[float, int]], m: int) -> List[int]:
        if len(cands) <= m:
            return [e for _, e in cands]
        chosen: List[int] = []
        for d, e in sorted(cands):
            if len(chosen) >= m:
                break
            if chosen:
                to_chosen = 1.0 - self.vecs[chosen] @ self.vecs[e]
                if bool((to_chosen < d).any()):      # closer to an already chosen neighbour than to the query
                    continue
            chosen.append(e)
        return chosen

    # ---- Alg. 1: insert ----------------------------------------------------------------------
    def add(self, vectors: np.ndarray) -> None:
        vectors = np.asarray(vectors, dtype=np.float32)
        vectors = vectors / np.maximum(np.linalg.norm(vectors, axis=1, keepdims=True), 1e-30)
        base = self.vecs.shape[0]
        self.vecs = np.concatenate([self.vecs, vectors])
        for i in range(base, base + vectors.shape[0]):
            level = int(-math.log(max(self.rng.random(), 1e-300)) * self.mult)
            self.links.append([[] for _ in range(level + 1)])
            if self.entry < 0:
                self.entry, self.max_level = i, level
                continue
            q = self.vecs[i]
            ep = [(float(self._dist(q, [self.entry])[0]), self.entry)]
            for lv in range(self.max_level, level, -1):
                ep = self._search_layer(q, ep, 1, lv)[:1]
            for lv in range(min(level, self.max_level), -1, -1):
                found = self._search_layer(q, ep, self.efc, lv)
                mmax = self.m0 if lv == 0 else self.m
                neigh = self._select(found, self.m)
                self.links[i][lv] = list(neigh)
                for n in neigh:                          # mutual links, shrunk with the same heuristic
                    ln = self.links[n][lv]
                    ln.append(i)
                    if len(ln) > mmax:
                        d = self._dist(self.vecs[n], ln).tolist()
                        self.links[n][lv] = self._select(list(zip(d, ln)), mmax)
                ep = found
            if level > self.max_level:
                self.entry, self.max_level = i, level

    # ---- Alg. 5: k-NN search -------------------------------------------------------------------
    def search(self, query: np.ndarray, k: int, ef: int = 100) -> Tuple[np.ndarray, np.ndarray]:
        """(ids [k], cosine similarity [k]) best first; similarity = 1 - distance like DenseIndex.search (:364-368)."""
        q = np.asarray(query, dtype=np.float32)
        q = q / max(float(np.linalg.norm(q)), 1e-30)
        if self.entry < 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
        ep = [(float(self._dist(q, [self.entry])[0]), self.entry)]
        for lv in range(self.max_level, 0, -1):
            ep = self._search_layer(q, ep, 1, lv)[:1]
        found = self._search_layer(q, ep, max(ef, k), 0)[:k]
        return (np.asarray([e for _, e in found], dtype=np.int64),
                np.asarray([1.0 - d for d, _ in found], dtype=np.float32))


def recall_at_k(approx_ids: np.ndarray, exact_ids: np.ndarray) -> float:
    """Mean fraction of the exact top-k ids that the approximate search returned (per query, then averaged)."""
    hits = [len(set(a.tolist()) & set(e.tolist())) / max(1, len(e)) for a, e in zip(approx_ids, exact_ids)]
    return float(np.mean(hits))


def recall_report(passages: np.ndarray, queries: np.ndarray, exact_ids: np.ndarray, k: int = 10,
                  efs: Sequence[int] = (10, 100)) -> dict:
    """Build the index over ``passages`` (ChromaDB defaults) and report recall@k against ``exact_ids`` [B, k]."""
    index = HnswCosine(passages.shape[1])
    index.add(passages)
    out = {"passages": int(passages.shape[0]), "queries": int(queries.shape[0]), "k": k, "M": index.m,
           "construction_ef": index.efc}
    for ef in efs:
        got = np.stack([np.pad(index.search(q, k, ef)[0], (0, k), constant_values=-1)[:k] for q in queries])
        out[f"recall@{k}_search_ef_{ef}"] = recall_at_k(got, exact_ids[:, :k])
    return out
