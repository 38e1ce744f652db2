"""Host-side evaluation metrics (scikit-learn), out of the accelerated scope.

Same functions, arguments and returned keys as shopformer/utils/metrics.py:18-77 so that the
reference scripts and the AUC parity tests can call them unchanged.
"""
from typing import Dict, Optional, Tuple

import numpy as np
from sklearn.metrics import (accuracy_score, average_precision_score, f1_score, precision_score, recall_score,
                             roc_auc_score, roc_curve)


def compute_auc_roc(labels: np.ndarray, scores: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """(auc, fpr, tpr); higher score = more anomalous."""
    labels, scores = np.asarray(labels), np.asarray(scores)
    fpr, tpr, _ = roc_curve(labels, scores)
    return float(roc_auc_score(labels, scores)), fpr, tpr


def compute_metrics(labels: np.ndarray, scores: np.ndarray, threshold: Optional[float] = None) -> Dict[str, float]:
    """AUC-ROC / AUC-PR plus accuracy, precision, recall, F1 at ``threshold`` (default: the
    Youden-J optimum of the ROC curve)."""
    labels, scores = np.asarray(labels), np.asarray(scores)
    auc_roc = roc_auc_score(labels, scores)
    auc_pr = average_precision_score(labels, scores)
    if threshold is None:
        fpr, tpr, thr = roc_curve(labels, scores)
        threshold = thr[int(np.argmax(tpr - fpr))]
    pred = (scores >= threshold).astype(int)
    return {
        "auc_roc": float(auc_roc),
        "auc_pr": float(auc_pr),
        "accuracy": float(accuracy_score(labels, pred)),
        "precision": float(precision_score(labels, pred, zero_division=0)),
        "recall": float(recall_score(labels, pred, zero_division=0)),
        "f1": float(f1_score(labels, pred, zero_division=0)),
        "threshold": float(threshold),
    }
