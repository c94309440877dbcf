"""CPU: the evaluation path and the data-side oracles (8f #2, #3) against vectors frozen from the
UNMODIFIED reference (oracle/make_golden_eval.py): `evaluate_ade_fde`, `evaluate_precision_recall`,
`Manifold.compute_inside`, `BaseDataset.ImageFeatures_small`."""
import os
import types

import numpy as np
import pytest
import torch

import data_oracle as DO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ev():
    z = np.load(os.path.join(GOLD, "evaluation.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def crops():
    z = np.load(os.path.join(GOLD, "scene_crop.npz"))
    return {k: z[k] for k in z.files}


def eval_ds(ev):
    sse = [tuple(int(x) for x in r) for r in ev["batch/seq_start_end"]]
    return types.SimpleNamespace(
        pred_traj=torch.from_numpy(ev["batch/gt_xy"]).permute(1, 0, 2).contiguous(),
        obs_traj=torch.from_numpy(ev["batch/in_xy"]).permute(1, 0, 2).contiguous(),
        seq_start_end=sse, scene_list=["synthetic_gofp"] * len(sse), dataset_name="synthetic_gofp")


def test_evaluate_ade_fde_matches_reference(ev):
    from mggan.evaluation import evaluate_ade_fde
    K = int(ev["meta/K"])
    got = evaluate_ade_fde(eval_ds(ev), ev["pred/abs"], list(range(1, K + 1)))
    keys = [k[len("ade_fde/"):] for k in ev if k.startswith("ade_fde/")]
    assert len(keys) == 3 * K and set(keys) == set(got)
    for k in keys:
        assert abs(got[k] - float(ev["ade_fde/" + k])) <= 1e-6 * max(1.0, abs(float(ev["ade_fde/" + k]))), k


@pytest.mark.parametrize("name", ["pred", "near"])
def test_precision_recall_matches_reference(ev, name):
    from mggan.evaluation import evaluate_precision_recall
    K = int(ev["meta/K"])
    got = evaluate_precision_recall(eval_ds(ev), ev[name + "/abs"], float(ev["meta/radius"]), list(range(1, K + 1)))
    keys = [k.split("/", 1)[1] for k in ev if k.startswith(f"pr_{name}/")]
    assert len(keys) == K + 1 and set(keys) == set(got)
    for k in keys:
        assert got[k] == pytest.approx(float(ev[f"pr_{name}/{k}"]), abs=1e-12), k


def test_manifold_inside_matches_reference(ev):
    from mggan.manifold import Manifold
    want = ev["inside/mask"]
    got = Manifold(ev["inside/manifold"], float(ev["meta/radius"])).compute_inside(ev["inside/tests"])
    assert np.array_equal(got, want)
    assert np.array_equal(DO.tube_inside(ev["inside/manifold"], ev["inside/tests"], float(ev["meta/radius"])), want)
    assert Manifold(ev["inside/manifold"], 1.0).compute_inside(ev["inside/tests"][:0]).shape == (0,)


def test_crop_oracle_matches_reference(crops):
    imgs = [crops[f"crop/image{i}"] for i in range(int(crops["crop/n_images"]))]
    feats = crops["crop/features"]
    assert len(feats) >= 30
    for j in range(len(feats)):
        i = int(crops["crop/image_id"][j])
        got = DO.image_features_small(imgs[i], crops["crop/last_xy"][j], float(crops["crop/scaling_small"][j]))
        assert np.array_equal(got, feats[j]), j
    # the fixture covers crops fully outside the image (all -1) and partially outside
    rgb = feats[:, :3]
    assert (rgb.reshape(len(feats), -1) == -1.0).all(1).any()
    assert ((rgb == -1.0).reshape(len(feats), -1).mean(1) > 0.2).sum() >= 4
