"""ADE / FDE evaluation over a dataset of predictions (reference: mggan/evaluation.py:14-78).

`evaluate_ade_fde` implements the reference's *intent*: as written the reference passes
`(…, None, "raw")` positionally into `(…, mode, mode_thresh)` and raises TypeError
(SURVEY.md 8a a17); the intended call is `mode="raw"` and that is what happens here.
Precision / Recall (manifold test, evaluation.py:101-156) is a "next" row of the scope table.
"""
from collections import defaultdict

import numpy as np
import torch

from mggan.metrics import compute_metrics_from_batch
from mggan.utils import to_numpy


def adjust_seq_start_end_for_mask(seq_start_end, remove_mask):
    """Scene ranges after dropping the agents flagged in `remove_mask` (reference :14-27)."""
    assert seq_start_end[-1][1] == len(remove_mask)
    offsets = [0] + np.cumsum(remove_mask).tolist()
    new_seq = [(s - offsets[s], e - offsets[e]) for s, e in seq_start_end]
    assert new_seq[-1][1] == np.sum(~remove_mask)
    return new_seq


def evaluate_ade_fde(eval_ds, preds, n_preds_list):
    """preds (pred_len, k, N_total, 2) numpy -> {"ADE k=i", "FDE k=i", "Mode k=i"}."""
    gt_trajs = to_numpy(eval_ds.pred_traj)                      # (N_total, pred_len, 2)
    pred_mask = np.isnan(gt_trajs).any(-1).any(-1)
    start_end = adjust_seq_start_end_for_mask(eval_ds.seq_start_end, pred_mask)
    gt_trajs = gt_trajs[~pred_mask]
    preds = preds[:, :, ~pred_mask]
    accum = defaultdict(lambda: np.zeros((2,)))
    for scene_idx, (start, end) in enumerate(start_end):
        if start == end:
            continue
        scaling = 1.0
        if eval_ds.dataset_name in ("stanford", "gofp"):        # pixel datasets
            scaling = 1.0 / eval_ds.images[eval_ds.scene_list[scene_idx]]["ratio"]
        gt = torch.from_numpy(gt_trajs[start:end]).transpose(0, 1) * scaling
        for n_preds in n_preds_list:
            m = compute_metrics_from_batch(torch.from_numpy(preds[:, :n_preds, start:end]) * scaling, gt,
                                           [[0, end - start]], mode="raw")
            for key, (value, count) in m.items():
                accum[f"{key} k={n_preds}"] += value, count
    return defaultdict(float, {key: value / count for key, (value, count) in accum.items()})
