"""Evaluation metrics (CPU / numpy, off the training hot path).  Same calling convention as the
reference: ``IMetric.__call__(prediction, target)`` on ``[N, user_sample_n]`` score matrices whose
positive item sits in column 0 (torchrec/metric/IMetric.py:17-26, Hit.py:20-23, NDCG.py:21-24,
MetricList.py:13-15).  ``LogLoss`` is added for the point-wise CTR models."""
from typing import Dict, List

import numpy as np


def get_pos_rank(prediction: np.ndarray, user_sample_n: int) -> np.ndarray:
    """Rank (1-based) of column 0 inside each group of ``user_sample_n`` scores."""
    scores = prediction.reshape(-1, user_sample_n)
    order = np.argsort(-scores, axis=1, kind="stable")
    return np.argmax(order == 0, axis=1) + 1


def get_pos_rank_torch(prediction, user_sample_n: int):
    """Device-side rank of column 0 (N3): 1 + number of candidates scoring strictly higher — the same answer as
    the stable argsort above, without moving the score matrix to the host."""
    scores = prediction.reshape(-1, user_sample_n)
    return 1 + (scores[:, 1:] > scores[:, :1]).sum(dim=1)


class IMetric:
    name = "metric"

    def __call__(self, prediction: np.ndarray, target: np.ndarray) -> float:
        raise NotImplementedError


class _RankMetric(IMetric):
    def __init__(self, user_sample_n: int, k: int):
        self.user_sample_n = user_sample_n
        self.k = k

    def from_rank(self, rank: np.ndarray) -> float:
        raise NotImplementedError

    def __call__(self, prediction, target):
        return self.from_rank(get_pos_rank(prediction, self.user_sample_n))


class Hit(_RankMetric):
    def __init__(self, user_sample_n: int, k: int):
        super().__init__(user_sample_n, k)
        self.name = f"hit@{k}"

    def from_rank(self, rank):
        return float((rank <= self.k).mean())


class NDCG(_RankMetric):
    def __init__(self, user_sample_n: int, k: int):
        super().__init__(user_sample_n, k)
        self.name = f"ndcg@{k}"

    def from_rank(self, rank):
        return float(((rank <= self.k) / np.log2(rank + 1)).mean())


class LogLoss(IMetric):
    """Binary cross-entropy of logits vs {0,1} targets."""
    name = "logloss"

    def __call__(self, prediction, target):
        z = prediction.astype(np.float64).reshape(-1)
        y = target.astype(np.float64).reshape(-1)
        return float(np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))


class MetricList:
    """Evaluates every metric; rank metrics with the same ``user_sample_n`` share one argsort."""

    def __init__(self, metrics: List[IMetric]):
        assert len(metrics) > 0
        self.metrics = metrics

    def rank_only(self) -> bool:
        return all(isinstance(m, _RankMetric) for m in self.metrics)

    def __call__(self, prediction, target) -> Dict[str, float]:
        ranks: Dict[int, np.ndarray] = {}
        out = {}
        on_device = not isinstance(prediction, np.ndarray)
        for m in self.metrics:
            if isinstance(m, _RankMetric):
                if m.user_sample_n not in ranks:
                    if on_device:  # torch tensor (any device): rank there, move only the [users] rank vector
                        ranks[m.user_sample_n] = get_pos_rank_torch(prediction, m.user_sample_n).cpu().numpy()
                    else:
                        ranks[m.user_sample_n] = get_pos_rank(prediction, m.user_sample_n)
                out[m.name] = m.from_rank(ranks[m.user_sample_n])
            else:
                out[m.name] = m(prediction, target)
        return out


def get_metric(metric_name: str) -> IMetric:
    """'ndcg@10' / 'hit@5' with the reference's fixed 99 negatives + 1 positive (metrics.py:13-15); 'logloss'."""
    name = metric_name.lower()
    if name == "logloss":
        return LogLoss()
    kind, k = name.split("@")
    cls = {"ndcg": NDCG, "hit": Hit}[kind]
    return cls(user_sample_n=100, k=int(k))
