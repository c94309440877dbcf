"""ADE / FDE evaluation over a dataset of predictions (reference: mggan/evaluation.py:14-78).

`evaluate_ade_fde` implements the reference's *intent*: as written the reference passes
`(…, None, "raw")` positionally into `(…, mode, mode_thresh)` and raises TypeError
(SURVEY.md 8a a17); the intended call is `mode="raw"` and that is what happens here.
`evaluate_precision_recall` is the manifold (growing-radius tube) test of evaluation.py:101-156.  Both host functions are
checked against outputs of the unmodified reference (tests/test_eval_crop_cpu.py); the `*_cuda` functions at the end of
the file return the same dicts with the arithmetic in `mggan_min_ade_fde` / `mggan_tube_inside`.
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


# ------------------------------------------------------------------------------------------------ device metrics
# Same results as the two host functions above, with the arithmetic in sm_100a kernels (csrc/data_eval.cu); the
# bookkeeping (masks, groups, descriptors) stays on the host like the reference's.  SURVEY.md 8f #3.
def _scene_scales(eval_ds, n_scenes):
    if eval_ds.dataset_name in ("stanford", "gofp"):        # pixel datasets (reference evaluation.py:60-61)
        return np.array([1.0 / eval_ds.images[eval_ds.scene_list[i]]["ratio"] for i in range(n_scenes)], np.float32)
    return None


def evaluate_ade_fde_cuda(eval_ds, preds, n_preds_list, device="cuda"):
    """`evaluate_ade_fde` on the device: preds (pred_len, k, N_total, 2) numpy or tensor.  One launch gives every
    k = 1 .. K at once (prefix minima over the samples)."""
    from mggan import kernels as K
    gt_trajs = to_numpy(eval_ds.pred_traj)
    pred_mask = np.isnan(gt_trajs).any(-1).any(-1)
    start_end = adjust_seq_start_end_for_mask(eval_ds.seq_start_end, pred_mask)
    keep = torch.from_numpy(np.where(~pred_mask)[0]).to(device)
    p = torch.as_tensor(preds).to(device=device, dtype=torch.float32)
    assert max(n_preds_list) <= p.shape[1], (max(n_preds_list), p.shape)
    p = p.index_select(2, keep).contiguous()
    gt = torch.from_numpy(gt_trajs[~pred_mask]).transpose(0, 1).contiguous().to(device)
    off = torch.tensor([start_end[0][0]] + [e for _, e in start_end], dtype=torch.int32, device=device)
    scales = _scene_scales(eval_ds, len(start_end))
    ade, fde, mode = K.min_ade_fde(p, gt, off, None if scales is None else torch.from_numpy(scales).to(device))
    ade, fde, mode = ade.sum(0).cpu().numpy(), fde.sum(0).cpu().numpy(), mode.sum(0).cpu().numpy()
    n, pred_len = gt.shape[1], gt.shape[0]
    out = defaultdict(float)
    for k in n_preds_list:
        out[f"FDE k={k}"] = fde[k - 1] / n
        out[f"ADE k={k}"] = ade[k - 1] / (pred_len * n)
        out[f"Mode k={k}"] = mode[k - 1] / n
    return out


def evaluate_precision_recall_cuda(eval_ds, all_preds, manifold_radius, n_preds_list, device="cuda"):
    """`evaluate_precision_recall` with every tube test of the sweep in ONE launch of `mggan_tube_inside`: trajectories go
    into one pool (ground-truth futures, then predictions agent-major), each test is a descriptor (test trajectory,
    manifold = a run of trajectory indices)."""
    from mggan import kernels as K
    gt_trajs = to_numpy(eval_ds.pred_traj).astype(np.float32)
    N, pred_len = gt_trajs.shape[:2]
    num_preds = max(n_preds_list)
    valid = np.where(~np.isnan(gt_trajs).any(-1).any(-1))[0]
    p = torch.as_tensor(all_preds).to(device=device, dtype=torch.float32)          # (pred_len, k, N, 2)
    k = p.shape[1]
    pool = torch.cat([torch.from_numpy(gt_trajs).to(device), p.permute(2, 1, 0, 3).reshape(N * k, pred_len, 2)])
    desc, man_list, segments = [], [], []                  # segments: (metric key, first descriptor, descriptors)
    for same_scene_indices in get_same_obs_indices(eval_ds):
        for same_ped_indices in zip(*same_scene_indices):
            peds = np.intersect1d(np.array(same_ped_indices), valid)
            if len(peds) == 0:
                continue
            rows = (N + peds[:, None] * k + np.arange(k)[None, :]).reshape(-1)     # cur_preds, agent-major
            gt_first = len(man_list)
            man_list.extend(peds.tolist())
            pr_first = len(man_list)
            man_list.extend(rows[:num_preds].tolist())
            tests = rows[:num_preds]
            segments.append(("Precision", len(desc), len(tests)))
            desc.extend((int(t), gt_first, len(peds)) for t in tests)
            for n_samples in n_preds_list:
                segments.append((f"Recall k={n_samples}", len(desc), len(peds)))
                desc.extend((int(t), pr_first, min(n_samples, len(rows))) for t in peds)
    if not desc:
        return defaultdict(float)
    radius = torch.from_numpy(np.linspace(manifold_radius / pred_len, manifold_radius, pred_len, endpoint=True)).to(device)
    inside = K.tube_inside(pool, radius, torch.tensor(desc, dtype=torch.int32, device=device),
                           torch.tensor(man_list, dtype=torch.int32, device=device)).cpu().numpy()
    accum = defaultdict(lambda: np.zeros((2,)))
    for key, first, count in segments:
        accum[key] += np.sum(inside[first:first + count]) / count, 1.0
    return defaultdict(float, {key: value / count for key, (value, count) in accum.items()})
