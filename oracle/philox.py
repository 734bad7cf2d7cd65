"""Philox4x32-10 with the curand / torch-CUDA counter layout (numpy).  TEST INFRASTRUCTURE ONLY.

The reference draws its dropout masks from whatever generator torch uses
(``nn.Dropout`` at ``rag_uq/router.py:78``); on CUDA that is Philox4x32-10
seeded as ``curand_init(seed, subsequence, offset)``:
    key     = (seed lo32, seed hi32)
    counter = (offset/4 lo32, offset/4 hi32, subsequence lo32, subsequence hi32)
and ``curand_uniform4`` maps each 32-bit output x to ``x * 2^-32 + 2^-33``
(float32).  This file restates that so the in-kernel masks can be checked
bit-for-bit without a GPU-side torch dependency.

``torch_dropout_geometry`` restates how torch's fused CUDA dropout kernel maps a
flat element index to (subsequence, draw index, lane) so the kernel's
``layout = torch`` mode can be checked against ``F.dropout`` on the GPU box.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for rnd in range(10):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
            if rnd != 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def uniform_from_bits(x: np.ndarray) -> np.ndarray:
    """curand_uniform: uint32 -> float32 in (0, 1]."""
    return x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def draw4(seed: int, subsequence: np.ndarray, offset4: np.ndarray) -> np.ndarray:
    """The 4 uint32 a thread with ``curand_init(seed, subsequence, 4*offset4)`` gets from curand4()."""
    subsequence = np.asarray(subsequence, dtype=np.uint64)
    offset4 = np.asarray(offset4, dtype=np.uint64)
    subsequence, offset4 = np.broadcast_arrays(subsequence, offset4)
    ctr = np.stack([offset4 & MASK32, offset4 >> np.uint64(32),
                    subsequence & MASK32, subsequence >> np.uint64(32)], axis=-1).astype(np.uint32)
    key = np.empty(ctr.shape[:-1] + (2,), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    return philox4x32_10(ctr, key)


def torch_dropout_geometry(n_elements: int, sm_count: int, max_threads_per_sm: int = 2048):
    """(grid, threads_total, offset_increment) torch's fused dropout uses for n float elements
    that are 16-byte aligned with n % 4 == 0 (vector width 4, block 256, unroll 4)."""
    block = 256
    grid = min(sm_count * (max_threads_per_sm // block), (n_elements + block - 1) // block)
    increment = ((n_elements - 1) // (block * grid * 4) + 1) * 4
    return grid, grid * block, increment


def keep_mask_torch_layout(n_elements: int, seed: int, offset: int, keep_prob: float,
                           sm_count: int) -> np.ndarray:
    """Keep mask (uint8, flat) torch-CUDA ``F.dropout`` would produce (vector-4 kernel).

    Element e belongs to thread ``(e // 4) % threads_total`` on its
    ``(e // 4) // threads_total``-th draw, lane ``e % 4``.
    """
    assert n_elements % 4 == 0 and offset % 4 == 0
    _, threads_total, _ = torch_dropout_geometry(n_elements, sm_count)
    quad = np.arange(n_elements // 4, dtype=np.uint64)
    bits = draw4(seed, quad % np.uint64(threads_total),
                 np.uint64(offset // 4) + quad // np.uint64(threads_total))
    u = uniform_from_bits(bits).reshape(-1)
    return (u < np.float32(keep_prob)).astype(np.uint8)
