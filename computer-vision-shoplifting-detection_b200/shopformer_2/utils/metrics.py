"""Host-side metrics of the variant-2 drop-in (scikit-learn; outside the accelerated scope).

Function names, arguments, returned keys and single-class fallbacks follow
shopformer_2/utils/metrics.py:21-205 so that the reference scripts run unchanged.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
from sklearn.metrics import (accuracy_score, average_precision_score, confusion_matrix, f1_score,
                             precision_recall_curve, precision_score, recall_score, roc_auc_score, roc_curve)


def compute_auc_roc(labels, scores) -> Tuple[float, np.ndarray, np.ndarray]:
    try:
        fpr, tpr, _ = roc_curve(labels, scores)
        return roc_auc_score(labels, scores), fpr, tpr
    except ValueError:                       # a single class present
        return 0.5, np.array([0, 1]), np.array([0, 1])


def compute_auc_pr(labels, scores) -> Tuple[float, np.ndarray, np.ndarray]:
    try:
        prec, rec, _ = precision_recall_curve(labels, scores)
        return average_precision_score(labels, scores), prec, rec
    except ValueError:
        return 0.0, np.array([0, 1]), np.array([1, 0])


def find_optimal_threshold(labels, scores, method: str = "youden") -> float:
    if method == "youden":
        fpr, tpr, thr = roc_curve(labels, scores)
        return thr[int(np.argmax(tpr - fpr))]
    if method == "f1":
        prec, rec, thr = precision_recall_curve(labels, scores)
        den = prec + rec
        f1 = np.where(den > 0, 2 * prec * rec / np.where(den > 0, den, 1), 0)
        return thr[int(np.argmax(f1[:-1]))]
    raise ValueError(f"Unknown method: {method}")


def compute_metrics(labels, scores, threshold: Optional[float] = None) -> Dict[str, float]:
    labels, scores = np.asarray(labels), np.asarray(scores)
    auc_roc, _, _ = compute_auc_roc(labels, scores)
    auc_pr, _, _ = compute_auc_pr(labels, scores)
    if threshold is None:
        try:
            threshold = find_optimal_threshold(labels, scores, "youden")
        except ValueError:
            threshold = float(np.median(scores))
    pred = (scores >= threshold).astype(int)
    try:
        tn, fp, fn, tp = confusion_matrix(labels, pred, labels=[0, 1]).ravel()
    except ValueError:
        tn = fp = fn = tp = 0
    return {
        "auc_roc": float(auc_roc), "auc_pr": float(auc_pr), "accuracy": float(accuracy_score(labels, pred)),
        "precision": float(precision_score(labels, pred, zero_division=0)),
        "recall": float(recall_score(labels, pred, zero_division=0)),
        "f1": float(f1_score(labels, pred, zero_division=0)), "threshold": float(threshold),
        "tp": int(tp), "fp": int(fp), "tn": int(tn), "fn": int(fn),
    }


def compute_video_level_metrics(video_scores: Dict[str, List[float]], video_labels: Dict[str, int],
                                aggregation: str = "max") -> Dict[str, float]:
    agg = {"max": np.max, "mean": np.mean, "percentile_95": lambda s: np.percentile(s, 95)}
    if aggregation not in agg:
        raise ValueError(f"Unknown aggregation: {aggregation}")
    vids = [v for v in video_scores if v in video_labels]
    scores = np.array([agg[aggregation](video_scores[v]) for v in vids])
    labels = np.array([video_labels[v] for v in vids])
    return compute_metrics(labels, scores)


def print_metrics(metrics: Dict[str, float], prefix: str = ""):
    head = f"{prefix} " if prefix else ""
    print(f"{head}AUC-ROC: {metrics['auc_roc']:.4f} | AUC-PR: {metrics['auc_pr']:.4f} | "
          f"Acc: {metrics['accuracy']:.4f} | P: {metrics['precision']:.4f} | R: {metrics['recall']:.4f} | "
          f"F1: {metrics['f1']:.4f} | thr: {metrics['threshold']:.4f}")
