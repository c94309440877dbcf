"""Scene-level displacement metrics (reference: mggan/metrics.py:6-68, :99-141).

minADE / minFDE are taken at SCENE level: the min over the k samples of the error summed over
all agents of the scene (so one sample index serves the whole scene), accumulated as
(sum, count) with counts pred_len * n and n.  Host-side, vectorised (no per-sample loop)."""
import numpy as np
import torch


def min_scene_error(error, seq_start_end):
    """error (k, b) -> sum over scenes of min_k sum_{agents in scene}."""
    total = 0.0
    for start, end in seq_start_end:
        total += error[:, start:end].sum(1).min(0)[0].item()
    return total


def displacement_error(pred_traj, pred_traj_gt, consider_ped=None, mode="sum"):
    """pred_traj, pred_traj_gt (T, b, 2) -> per-agent sum over time of the Euclidean error."""
    loss = (pred_traj_gt - pred_traj).pow(2).sum(-1).sqrt().sum(0)
    if consider_ped is not None:
        loss = loss * consider_ped
    return loss if mode == "raw" else loss.sum()


def final_displacement_error(pred_pos, pred_pos_gt, consider_ped=None, mode="sum"):
    loss = (pred_pos_gt - pred_pos).pow(2).sum(-1).sqrt()
    if consider_ped is not None:
        loss = loss * consider_ped
    return loss if mode == "raw" else loss.sum()


def compute_metrics_from_batch(preds, gt, sub_batches, mode="mean", mode_thresh=3.0):
    """preds (T, k, b, 2), gt (T, b, 2) -> {"FDE", "ADE", "Mode"} as value/count pairs
    (mode="raw") or their ratios (mode="mean")."""
    pred_len, k, b, _ = preds.shape
    err = (preds - gt[:, None]).pow(2).sum(-1).sqrt()          # (T, k, b)
    ades, fdes = err.sum(0), err[-1]
    metrics = {
        "FDE": np.array([min_scene_error(fdes, sub_batches), b]),
        "ADE": np.array([min_scene_error(ades, sub_batches), pred_len * b]),
        "Mode": np.array([(fdes.min(0)[0] < mode_thresh).float().sum().item(), b]),
    }
    if mode == "mean":
        return {key: (v / c) for key, (v, c) in metrics.items()}
    return metrics
