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


def get_same_obs_indices(eval_ds):
    """Groups of scenes that share their observations (multi-future datasets): list of groups, each a
    list of per-scene agent-index lists (reference :30-40)."""
    obs_trajs = to_numpy(eval_ds.obs_traj)
    groups = defaultdict(list)
    for scene_idx, (start, end) in enumerate(eval_ds.seq_start_end):
        key = (obs_trajs[start:end].tobytes(), eval_ds.scene_list[scene_idx])
        groups[key].append(list(range(start, end)))
    return list(groups.values())


def evaluate_precision_recall(eval_ds, all_preds, manifold_radius, n_preds_list, debug=False):
    """Precision: fraction of predictions inside the tube around the ground-truth futures that share
    the observation; Recall k=i: fraction of those futures inside the tube around the first i
    predictions (reference :101-156)."""
    from mggan.manifold import Manifold
    gt_trajs = to_numpy(eval_ds.pred_traj)
    num_preds = max(n_preds_list)
    pred_mask = np.isnan(gt_trajs).any(-1).any(-1)
    valid = np.where(~pred_mask)[0]
    preds = all_preds.transpose(2, 1, 0, 3)                    # (N, k, pred_len, 2)
    accum = defaultdict(lambda: np.zeros((2,)))
    for same_scene_indices in get_same_obs_indices(eval_ds):
        for same_ped_indices in zip(*same_scene_indices):
            same_ped_indices = np.intersect1d(np.array(same_ped_indices), valid)
            if len(same_ped_indices) == 0:
                continue
            gt_man_samples = gt_trajs[same_ped_indices]
            gt_man = Manifold(gt_man_samples, manifold_radius)
            cur_preds = preds[same_ped_indices].reshape(-1, *preds.shape[2:])
            accum["Precision"] += gt_man.compute_metric(cur_preds[:num_preds]), 1.0
            for n_samples in n_preds_list:
                pred_man = Manifold(cur_preds[:n_samples], manifold_radius)
                accum[f"Recall k={n_samples}"] += pred_man.compute_metric(gt_man_samples), 1.0
    return defaultdict(float, {key: value / count for key, (value, count) in accum.items()})
