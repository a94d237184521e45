"""Group-fairness metrics on top of the segmented-sort kernel — mirrors evaluation/metrics.py of the reference.

The GPU produces INTEGER Mann–Whitney / confusion counts for every (attribute, group) slot in one pass
(ops.group_auc_counts -> csrc/group_auc.cu); this module turns them into the reference's scores with the
reference's formulas and its quirks (evaluation/metrics.py:197-311, 340-356, 486-550):
  * `compute_auc` on [N,2] probabilities = macro one-vs-rest AUC = mean of the two column AUCs;
  * ES-AUC and the per-group AUC list skip the unknown group -1, ES-ACC / DPD / EOD / AOD do not;
  * a group containing a single class has no AUC: the reference prints and exit()s — here a ValueError.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops

GT0, EQ0, GT1, EQ1, TP, FP, TN, FN = range(8)


class GroupCounts:
    """Counts table [n_slots, 8] with slot helpers (slot 0 = overall; then (attribute, group+1))."""

    def __init__(self, table: np.ndarray, n_attr: int, max_groups: int):
        self.t = table.astype(np.int64)
        self.n_attr, self.max_groups = n_attr, max_groups

    def slot(self, attr_idx: Optional[int] = None, group: int = 0) -> np.ndarray:
        if attr_idx is None:
            return self.t[0]
        return self.t[1 + attr_idx * (self.max_groups + 1) + (group + 1)]

    def groups_present(self, attr_idx: int, include_unknown: bool):
        lo = -1 if include_unknown else 0
        return [g for g in range(lo, self.max_groups) if int(self.slot(attr_idx, g)[TP:FN + 1].sum()) > 0]


def group_counts(prob, y, attrs=None, max_groups: Optional[int] = None, device=None) -> GroupCounts:
    """Run the kernel. prob [N,2] (numpy or tensor), y [N], attrs [n_attr, N] with -1 = unknown."""
    if not torch.cuda.is_available():
        raise RuntimeError("fairfedmed_b200.metrics needs a CUDA device (no CPU fallback)")
    device = device or torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(np.asarray(prob) if not torch.is_tensor(prob) else prob).to(device=device, dtype=torch.float32)
    lab = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(device)
    a = None
    n_attr = 0
    if attrs is not None:
        a = torch.as_tensor(np.asarray(attrs) if not torch.is_tensor(attrs) else attrs).to(device)
        n_attr = a.shape[0]
        if max_groups is None:
            max_groups = int(a.max().item()) + 1 if a.numel() else 1
    max_groups = max(1, max_groups or 1)
    table = ops.group_auc_counts(p, lab, a, max_groups).cpu().numpy()
    return GroupCounts(table, n_attr, max_groups)


def _auc_from_row(row: np.ndarray) -> float:
    n_pos1 = int(row[TP] + row[FN])     # label == 1
    n_pos0 = int(row[TN] + row[FP])     # label == 0
    if n_pos1 == 0 or n_pos0 == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    denom = float(n_pos1) * float(n_pos0)
    auc1 = (float(row[GT1]) + 0.5 * float(row[EQ1])) / denom
    auc0 = (float(row[GT0]) + 0.5 * float(row[EQ0])) / denom
    return float(np.mean([auc0, auc1]))


def _binary_auc_from_row(row: np.ndarray) -> float:
    n_pos1 = int(row[TP] + row[FN])
    n_pos0 = int(row[TN] + row[FP])
    if n_pos1 == 0 or n_pos0 == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return (float(row[GT1]) + 0.5 * float(row[EQ1])) / (float(n_pos1) * float(n_pos0))


def compute_auc(pred_prob, y, num_classes=2):
    """evaluation/metrics.py:340-356."""
    if torch.is_tensor(pred_prob):
        pred_prob = pred_prob.detach()
    pp = pred_prob if torch.is_tensor(pred_prob) else np.asarray(pred_prob)
    if num_classes != 2:
        raise NotImplementedError("the FairLoRA path is binary (num_classes = 2)")
    if tuple(pp.shape) == tuple(np.shape(y) if not torch.is_tensor(y) else y.shape):
        # one score per sample: rank statistics of column 1 only
        p1 = torch.as_tensor(pp, dtype=torch.float32).reshape(-1)
        two = torch.stack([torch.zeros_like(p1), p1], dim=1)
        return _binary_auc_from_row(group_counts(two, y).slot())
    return _auc_from_row(group_counts(pp, y).slot())


def accuracy(output, target, topk=(1,)):
    """evaluation/metrics.py:313-338 for [N,2] probabilities (top-1)."""
    c = group_counts(output, target).slot()
    return float(c[TP] + c[TN]) / float(c[TP:FN + 1].sum())


def _acc(row):
    n = float(row[TP:FN + 1].sum())
    return float(row[TP] + row[TN]) / n


def _rates(row):
    n = float(row[TP:FN + 1].sum())
    pos, neg = float(row[TP] + row[FN]), float(row[TN] + row[FP])
    sel = float(row[TP] + row[FP]) / n if n else float("nan")
    tpr = float(row[TP]) / pos if pos else float("nan")
    fpr = float(row[FP]) / neg if neg else float("nan")
    return sel, tpr, fpr


def equity_scaled_accuracy_from(counts: GroupCounts, attr_idx: int, alpha=1.0) -> float:
    overall = _acc(counts.slot())
    gap = sum(abs(_acc(counts.slot(attr_idx, g)) - overall) for g in counts.groups_present(attr_idx, True))
    return overall / (alpha * gap + 1)


def equity_scaled_auc_from(counts: GroupCounts, attr_idx: int, alpha=1.0) -> float:
    overall = _auc_from_row(counts.slot())
    gap = sum(abs(_auc_from_row(counts.slot(attr_idx, g)) - overall) for g in counts.groups_present(attr_idx, False))
    return overall / (alpha * gap + 1)


def equity_scaled_accuracy(output, target, attrs, alpha=1.0):
    """evaluation/metrics.py:486-511 (single attribute vector `attrs` [N])."""
    a = np.asarray(attrs.cpu() if torch.is_tensor(attrs) else attrs).reshape(1, -1)
    return equity_scaled_accuracy_from(group_counts(output, target, a), 0, alpha)


def equity_scaled_AUC(output, target, attrs, alpha=1.0, num_classes=2):
    """evaluation/metrics.py:513-547."""
    a = np.asarray(attrs.cpu() if torch.is_tensor(attrs) else attrs).reshape(1, -1)
    return equity_scaled_auc_from(group_counts(output, target, a), 0, alpha)


def compute_between_group_disparity(auc_list, overall_auc):
    """evaluation/metrics.py:549-550."""
    return np.std(auc_list) / overall_auc, (np.max(auc_list) - np.min(auc_list)) / overall_auc


def _dpd(counts, a):
    sel = [_rates(counts.slot(a, g))[0] for g in counts.groups_present(a, True)]
    return float(np.max(sel) - np.min(sel))


def _eod(counts, a):
    r = [_rates(counts.slot(a, g)) for g in counts.groups_present(a, True)]
    tpr, fpr = [x[1] for x in r], [x[2] for x in r]
    return float(max(np.nanmax(tpr) - np.nanmin(tpr), np.nanmax(fpr) - np.nanmin(fpr)))


def _aod(counts, a):
    tot = counts.slot()
    vals = []
    for g in counts.groups_present(a, True):
        row = counts.slot(a, g)
        rest = tot - row
        _, tpr_p, fpr_p = _rates(row)
        _, tpr_u, fpr_u = _rates(rest)
        tpr_p, fpr_p, tpr_u, fpr_u = (0.0 if np.isnan(v) else v for v in (tpr_p, fpr_p, tpr_u, fpr_u))
        vals.append(abs(((fpr_u - fpr_p) + (tpr_u - tpr_p)) / 2))
    return sum(vals) / max(len(vals), 1)


def evalute_comprehensive_perf_scores(preds, gts, attrs=None, num_classes=2):
    """Drop-in for evaluation/metrics.py:197-311 (binary, [N,2] probabilities): ONE kernel pass for all attributes.

    Returns overall_acc, esaccs_by_attrs, overall_auc, esaucs_by_attrs, aucs_by_attrs, dpds, eods, aods,
    between_group_disparity — same order and container types as the reference.  DPD / EOD / AOD follow the public
    fairlearn / aif360 definitions (un-vendored upstream: "parity unpinned")."""
    if num_classes != 2:
        raise NotImplementedError("the FairLoRA path is binary (num_classes = 2)")
    counts = group_counts(preds, gts, attrs)
    overall_acc = _acc(counts.slot())
    overall_auc = _auc_from_row(counts.slot())
    esaccs, esaucs, aucs_by_attrs, dpds, eods, aods, disp = [], [], [], [], [], [], []
    for a in range(counts.n_attr):
        esaccs.append(equity_scaled_accuracy_from(counts, a))
        esaucs.append(equity_scaled_auc_from(counts, a))
        g_aucs = [_auc_from_row(counts.slot(a, g)) for g in counts.groups_present(a, False)]
        aucs_by_attrs.append(np.array(g_aucs))
        disp.append(list(compute_between_group_disparity(g_aucs, overall_auc)))
        dpds.append(_dpd(counts, a))
        eods.append(_eod(counts, a))
        aods.append(_aod(counts, a))
    return (overall_acc, np.array(esaccs), overall_auc, np.array(esaucs), aucs_by_attrs, np.array(dpds),
            np.array(eods), aods, np.array(disp))
