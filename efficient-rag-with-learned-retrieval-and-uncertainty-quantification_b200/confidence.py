"""Uncertainty result types and the MC aggregation arithmetic.

``ConfidenceResult`` is the reference's result record (rag_uq/confidence.py:46-55).
``RouterUncertainty`` carries what the MC-Dropout kernel produces and converts to that record
with the reference's own scaling (``uncertainty = min(1, variance / 2)``,
``confidence = 1 - uncertainty``, confidence.py:258-264).

``MCDropoutConfidence`` keeps the reference class's surface (confidence.py:69-272).  Its
sampling loop is bound by an external LLM and stays on the host; only the aggregation
(centroid / distance / std, :195-202; consensus = argmin, :247-250) is arithmetic, and it is
the same arithmetic the kernel applies to the router's gate vectors.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)


@dataclass
class ConfidenceResult:
    """Result from confidence estimation (rag_uq/confidence.py:46-55)."""
    answers: List[str]
    consensus_answer: str
    uncertainty_score: float
    confidence: float
    embedding_variance: Optional[float] = None
    lexical_diversity: Optional[float] = None
    metadata: Dict[str, Any] = field(default_factory=dict)


@dataclass
class RouterUncertainty:
    """MC-Dropout statistics of the router gate over T samples (device tensors).

    mean_gate / std_gate / mean_fused / std_fused: [B, P]; std is the population std (ddof 0)
    like ``distances.std()`` in the reference.  variance [B]: std over the T samples of the L2
    distance between a sample's gate vector and the centroid.  consensus [B]: index of the
    sample closest to the centroid.
    """
    mean_gate: torch.Tensor
    std_gate: torch.Tensor
    mean_fused: torch.Tensor
    std_fused: torch.Tensor
    variance: torch.Tensor
    consensus: torch.Tensor
    n_samples: int
    masks: Optional[torch.Tensor] = None
    gates: Optional[torch.Tensor] = None

    @property
    def uncertainty(self) -> torch.Tensor:
        return torch.clamp(self.variance / 2.0, max=1.0)     # confidence.py:258

    @property
    def confidence(self) -> torch.Tensor:
        return 1.0 - self.uncertainty                          # confidence.py:264

    def to_confidence_result(self, query: int = 0, doc_ids: Optional[List[str]] = None) -> ConfidenceResult:
        """The reference's record for one query; 'answers' are the candidate ids ranked by mean fused score."""
        order = torch.argsort(self.mean_fused[query], descending=True).tolist()
        names = [doc_ids[i] if doc_ids is not None else str(i) for i in order]
        var = float(self.variance[query])
        unc = min(1.0, var / 2.0)
        return ConfidenceResult(
            answers=names, consensus_answer=names[0] if names else "", uncertainty_score=unc, confidence=1.0 - unc,
            embedding_variance=var, lexical_diversity=None,
            metadata={"n_samples": self.n_samples, "consensus_sample": int(self.consensus[query]),
                      "source": "router-mc-dropout"})


def embedding_variance(embeddings: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """(std of distances to the centroid, centroid, distances) - confidence.py:195-202."""
    centroid = embeddings.mean(axis=0)
    distances = np.linalg.norm(embeddings - centroid, axis=1)
    return float(distances.std()), centroid, distances


def lexical_diversity(answers: List[str]) -> float:
    """Type/token ratio over all answers, 1.0 when there are no tokens - confidence.py:164-175."""
    tokens = [tok for a in answers for tok in a.lower().split()]
    return len(set(tokens)) / len(tokens) if tokens else 1.0


class MCDropoutConfidence:
    """Host-side mirror of rag_uq/confidence.py:69-272 (LLM re-prompting; not a kernel target).

    ``encoder`` may be any object with ``encode(list[str]) -> ndarray``; when omitted the class
    tries ``sentence_transformers`` like the reference and otherwise works without embeddings.
    """

    def __init__(self, llm_client, n_samples: int = 10, embedding_model: str = "all-MiniLM-L6-v2",
                 temperature_range: Tuple[float, float] = (0.5, 1.2), top_p_range: Tuple[float, float] = (0.8, 0.95),
                 max_tokens: int = 100, encoder: Any = None):
        self.llm = llm_client
        self.n_samples = n_samples
        self.temperature_range = temperature_range
        self.top_p_range = top_p_range
        self.max_tokens = max_tokens
        self.encoder = encoder
        if self.encoder is None:
            try:
                from sentence_transformers import SentenceTransformer
                self.encoder = SentenceTransformer(embedding_model)
            except ImportError:
                logger.warning("Sentence encoder not available")

    def _sample_parameters(self) -> Dict[str, float]:
        return {"temperature": np.random.uniform(*self.temperature_range),
                "top_p": np.random.uniform(*self.top_p_range)}

    def _generate_sample(self, prompt: str, context: str, question: str, model: str = "llama3.2:3b") -> str:
        params = self._sample_parameters()
        full_prompt = f"{prompt}\n\nContext: {context}\n\nQuestion: {question}\n\nAnswer:"
        try:
            response = self.llm.generate(model=model, prompt=full_prompt,
                                         options={"temperature": params["temperature"], "top_p": params["top_p"],
                                                  "num_predict": self.max_tokens})
            return response.get("response", "").strip()
        except Exception as exc:  # same soft-fail default as the reference (:160-162)
            logger.error(f"LLM generation failed: {exc}")
            return ""

    _compute_lexical_diversity = staticmethod(lexical_diversity)

    def _compute_embedding_variance(self, answers: List[str]):
        valid = [a for a in answers if a.strip()] if answers else []
        if self.encoder is None or not valid:
            return 1.0, np.array([]), np.array([])
        emb = np.asarray(self.encoder.encode(valid))
        var, centroid, _ = embedding_variance(emb)
        return var, centroid, emb

    def get_confidence_interval(self, prompt: str, context: str, question: str,
                                model: str = "llama3.2:3b") -> ConfidenceResult:
        answers = [a for a in (self._generate_sample(prompt, context, question, model) for _ in range(self.n_samples)) if a]
        if not answers:
            return ConfidenceResult([], "", 1.0, 0.0, metadata={"error": "No valid answers generated"})
        diversity = lexical_diversity(answers)
        variance, centroid, emb = self._compute_embedding_variance(answers)
        if len(emb) > 0:
            nearest = int(np.argmin(np.linalg.norm(emb - centroid, axis=1)))
            consensus = [a for a in answers if a.strip()][nearest]
        else:
            from collections import Counter
            consensus = Counter(answers).most_common(1)[0][0]
        unc = min(1.0, variance / 2.0)
        return ConfidenceResult(answers, consensus, unc, 1.0 - unc, variance, diversity,
                                {"n_samples": len(answers), "temperature_range": self.temperature_range,
                                 "top_p_range": self.top_p_range})
