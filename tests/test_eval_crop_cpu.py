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


def test_device_metric_host_logic_matches_reference(ev, monkeypatch):
    """The descriptor / offset construction and the reductions of `evaluate_*_cuda` (host logic), with the two kernels
    replaced by straightforward torch / numpy stand-ins; the kernels themselves are checked on the GPU
    (tests/test_gpu_zd_data_eval.py)."""
    from mggan import evaluation as E
    from mggan import kernels as K

    def fake_tube(traj, radius, desc, man_list):
        traj, r, out = traj.numpy(), radius.numpy(), []
        for t, f, c in desc.tolist():
            out.append(bool(DO.tube_inside(traj[man_list[f:f + c].numpy()], traj[t][None], float(r[-1]))[0]))
        return torch.tensor(out)

    def fake_min(preds, gt, off, scale=None, mode_thresh=3.0):
        T, Kk, n, _ = preds.shape
        S = off.numel() - 1
        err = (preds - gt[:, None]).pow(2).sum(-1).sqrt().double()
        ade, fde = torch.zeros(S, Kk, dtype=torch.float64), torch.zeros(S, Kk, dtype=torch.float64)
        mode = torch.zeros(S, Kk, dtype=torch.int32)
        for s in range(S):
            a, b = int(off[s]), int(off[s + 1])
            ade[s] = torch.cummin(err[:, :, a:b].sum(0).sum(1), 0).values
            fde[s] = torch.cummin(err[-1, :, a:b].sum(1), 0).values
            mode[s] = (torch.cummin(err[-1, :, a:b], 0).values < mode_thresh).sum(1).int()
        return ade, fde, mode

    monkeypatch.setattr(K, "tube_inside", fake_tube)
    monkeypatch.setattr(K, "min_ade_fde", fake_min)
    ds, Kk = eval_ds(ev), int(ev["meta/K"])
    nl = list(range(1, Kk + 1))
    for name in ("pred", "near"):
        got = E.evaluate_precision_recall_cuda(ds, ev[name + "/abs"], float(ev["meta/radius"]), nl, device="cpu")
        assert len(got) == Kk + 1
        for k in got:
            assert got[k] == pytest.approx(float(ev[f"pr_{name}/{k}"]), abs=1e-12), (name, k)
    got = E.evaluate_ade_fde_cuda(ds, ev["pred/abs"], nl, device="cpu")
    assert len(got) == 3 * Kk
    for k in got:
        assert got[k] == pytest.approx(float(ev["ade_fde/" + k]), rel=1e-5), k


def test_synthetic_scene_image_modes():
    """Dataset side of the resident-image path: host-cut features equal the crop oracle, the resident mode carries
    image ids that index the dataset's image list, multi-future replicas share their scene's image."""
    from mggan.data_utils.data_loaders import get_dataloader
    from mggan.synthetic import SCALING_SMALL, make_image_batch
    host, images = make_image_batch([3, 5, 1], seed=4)
    res, _ = make_image_batch([3, 5, 1], seed=4, resident=True)
    assert "features" not in res and res["image_ids"].dtype == np.int32
    assert res["image_ids"].tolist() == [0] * 3 + [1] * 5 + [2]
    last = host["in_xy"][-1]
    for a in range(9):
        want = DO.image_features_small(images[res["image_ids"][a]], last[a], SCALING_SMALL)
        assert np.array_equal(want, host["features"][a])
    dl_h = get_dataloader("synthetic_gofp", batch_size=3, num_scenes=5, images="host_crop")
    dl_r = get_dataloader("synthetic_gofp", batch_size=3, num_scenes=5, images="resident")
    imgs = dl_r.dataset.scene_image_list()
    assert len(imgs) == 5
    for bh, br in zip(dl_h, dl_r):
        assert torch.equal(bh["in_xy"], br["in_xy"]) and bh["seq_start_end"] == br["seq_start_end"]
        ids = br["image_ids"].numpy()
        assert ids.min() >= 0 and ids.max() < 5 and "features" not in br
        for a in range(0, len(ids), 5):
            want = DO.image_features_small(imgs[ids[a]], bh["in_xy"][-1, a].numpy(), SCALING_SMALL)
            assert np.array_equal(want, bh["features"][a].numpy())


def test_oracle_reproduces_the_fixture_predictions(ev):
    """The evaluation fixture's predictions came from the reference's `predict` (eval mode, injected noise and PM-Network
    draws): the oracle generator reproduces them, so the GPU test built on the fixture asks the right thing."""
    import mggan_oracle as O
    sd = {k[2:]: torch.from_numpy(v) for k, v in ev.items() if k.startswith("G/")}
    sse = [tuple(int(x) for x in r) for r in ev["batch/seq_start_end"]]
    with torch.no_grad():
        (_, ab), logits, _ = O.generator_forward(
            sd, 4, torch.from_numpy(ev["batch/in_xy"]), torch.from_numpy(ev["batch/in_dxdy"]), sse,
            torch.from_numpy(ev["pred/noise"]), False, None, int(ev["meta/K"]), None, torch.from_numpy(ev["pred/idx"]),
            training=False)
    assert float((ab - torch.from_numpy(ev["pred/abs"])).abs().max()) <= 1e-5 * float(np.abs(ev["pred/abs"]).max())
    assert float((torch.softmax(logits, 1) - torch.from_numpy(ev["pred/probs"])).abs().max()) <= 1e-6
