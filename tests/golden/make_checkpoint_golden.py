"""A genuine reference router checkpoint + the reference's own outputs for it (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_checkpoint_golden.py

The LIVE ``rag_uq.router.RouterTrainer`` from /root/reference trains a ``RetrievalRouter`` for a few steps on
synthetic score pairs (so the running statistics, the optimizer state and the loss history are real) and writes the
checkpoint with its own ``save_checkpoint`` (router.py:499-508): a ``torch.save`` of ``model_state_dict``,
``optimizer_state_dict``, the pickled ``RouterConfig`` instance, ``train_losses`` and ``val_losses``.  Next to it go the
inputs and what the reference module answers for them after ``load_checkpoint`` - in eval mode, both right after loading
(``stats_initialized`` False: call-wide statistics) and with the running statistics armed.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")
from rag_uq.router import RetrievalRouter, RouterConfig, RouterTrainer  # noqa: E402

torch.manual_seed(2026)
router = RetrievalRouter(RouterConfig(hidden_dim=32, dropout=0.2))
trainer = RouterTrainer(router, learning_rate=1e-2)
g = torch.Generator().manual_seed(5)
router.train()
for step in range(6):                                   # a few real optimisation steps: EMA statistics + Adam state
    b = torch.rand(8, 20, generator=g) * 12.0
    d = torch.rand(8, 20, generator=g) * 1.4 - 0.2
    rel = (torch.rand(8, 20, generator=g) > 0.8).float()
    trainer.optimizer.zero_grad()
    w = router(b, d)
    loss = trainer.loss_fn(w * d + (1 - w) * b, rel)
    loss.backward()
    trainer.optimizer.step()
    trainer.train_losses.append(float(loss))
path = HERE / "router_checkpoint.pt"
trainer.save_checkpoint(str(path))

fresh = RetrievalRouter(RouterConfig(hidden_dim=32, dropout=0.2))
# torch >= 2.6 unpickles with weights_only=True by default, which rejects the pickled RouterConfig: the reference's own
# load_checkpoint (router.py:510-517) only works with the class allow-listed
with torch.serialization.safe_globals([RouterConfig]):
    RouterTrainer(fresh).load_checkpoint(str(path))
fresh.eval()
b = torch.rand(5, 20, generator=g) * 12.0
d = torch.rand(5, 20, generator=g) * 1.4 - 0.2
with torch.no_grad():
    gate_call = fresh(b, d).numpy().copy()               # stats_initialized is NOT part of the state dict: False after loading
    vals, idx = fresh.hybrid_rerank(b, d, top_k=10)
    fresh.stats_initialized = True
    gate_running = fresh(b, d).numpy().copy()
np.savez_compressed(HERE / "router_checkpoint_expected.npz", bm25=b.numpy(), dense=d.numpy(), gate_call=gate_call,
                    rerank_vals=vals.numpy(), rerank_idx=idx.numpy(), gate_running=gate_running,
                    running_stats=np.array([float(fresh.bm25_mean), float(fresh.bm25_std), float(fresh.dense_mean),
                                            float(fresh.dense_std)], dtype=np.float32))
print("wrote", path, path.stat().st_size, "bytes")
