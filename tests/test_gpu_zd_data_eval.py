"""GPU parity of the data-side / evaluation kernels (csrc/data_eval.cu) against
vectors frozen from the UNMODIFIED reference (tests/golden/scene_crop.npz, evaluation.npz) and against the numpy oracle
(oracle/data_oracle.py) on seeded inputs.  Crops and tube decisions are bit-exact; ADE / FDE within 1e-2 (north_star).
(The file name sorts after the training-step parity tests on purpose: those ran on a B200 before this file existed.)"""
import os
import types

import numpy as np
import pytest
import torch

import data_oracle as DO

pytestmark = pytest.mark.gpu
DEV = "cuda"
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


# ------------------------------------------------------------------------------------------------ crop
def test_scene_crop_matches_reference_bit_exact(crops):
    from mggan.data_utils.scene_images import SceneImageStore
    imgs = [crops[f"crop/image{i}"] for i in range(int(crops["crop/n_images"]))]
    scal = [float(crops["crop/scaling_small"][list(crops["crop/image_id"]).index(i)]) for i in range(len(imgs))]
    store = SceneImageStore(imgs, scal, device=DEV)
    got = store.crop(crops["crop/image_id"], torch.from_numpy(crops["crop/last_xy"]).to(DEV)).cpu().numpy()
    assert got.shape == crops["crop/features"].shape
    assert np.array_equal(got, crops["crop/features"])


@pytest.mark.parametrize("n", [0, 1, 257, 5000])
def test_scene_crop_matches_oracle_ragged(n):
    from mggan.data_utils.scene_images import SceneImageStore
    rng = np.random.default_rng(n + 1)
    imgs = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for h, w in ((40, 70), (128, 96), (1, 1), (33, 200))]
    scal = [0.5, 0.7, 1.2, 0.5]
    store = SceneImageStore(imgs, scal, device=DEV)
    ids = rng.integers(0, len(imgs), size=n).astype(np.int32)
    xy = rng.uniform(-12.0, 80.0, size=(n, 2)).astype(np.float32)
    if n > 2:
        ids[1] = -1                         # no image: RGB channels are -1
        ids[2] = len(imgs)
    got = store.crop(ids, torch.from_numpy(xy).to(DEV)).cpu().numpy()
    assert got.shape == (n, 4, 33, 33)
    for j in range(n):
        if 0 <= ids[j] < len(imgs):
            want = DO.image_features_small(imgs[ids[j]], xy[j], scal[ids[j]])
        else:
            want = DO.image_features_small(np.zeros((1, 1, 3), np.uint8), np.array([-100.0, -100.0], np.float32), 1.0)
        assert np.array_equal(got[j], want), j


def test_cropped_features_feed_the_scene_cnn(crops):
    """Crops cut on the device are a drop-in for the host-built `features` tensor of a batch."""
    from mggan.data_utils.scene_images import SceneImageStore
    from mggan.model.modules.cnn import AttentionGlobal
    imgs = [crops[f"crop/image{i}"] for i in range(int(crops["crop/n_images"]))]
    scal = [float(crops["crop/scaling_small"][list(crops["crop/image_id"]).index(i)]) for i in range(len(imgs))]
    store = SceneImageStore(imgs, scal, device=DEV)
    torch.manual_seed(0)
    net = AttentionGlobal(channels_cnn=16).to(DEV)
    net.eval()
    with torch.no_grad():
        a = net(store.crop(crops["crop/image_id"], torch.from_numpy(crops["crop/last_xy"]).to(DEV)))
        b = net(torch.from_numpy(crops["crop/features"]).to(DEV))
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)      # identical inputs; equal up to the order of any atomic sums


# ------------------------------------------------------------------------------------------------ tube test
def test_tube_inside_matches_reference_mask(ev):
    from mggan import kernels as K
    man, tests = ev["inside/manifold"], ev["inside/tests"]
    pool = torch.from_numpy(np.concatenate([man, tests])).to(DEV)
    T, r = man.shape[1], float(ev["meta/radius"])
    radius = torch.from_numpy(np.linspace(r / T, r, T)).to(DEV)
    desc = torch.tensor([(len(man) + i, 0, len(man)) for i in range(len(tests))], dtype=torch.int32, device=DEV)
    lst = torch.arange(len(man), dtype=torch.int32, device=DEV)
    got = K.tube_inside(pool, radius, desc, lst).cpu().numpy()
    assert np.array_equal(got, ev["inside/mask"])


@pytest.mark.parametrize("m,n", [(1, 1), (3, 40), (70, 33), (0, 5)])
def test_tube_inside_matches_oracle(m, n):
    from mggan import kernels as K
    rng = np.random.default_rng(10 * m + n)
    T, r = 12, 0.9
    man = rng.normal(0, 1.0, size=(m, T, 2)).astype(np.float32)
    tests = rng.normal(0, 0.6, size=(n, T, 2)).astype(np.float32)
    if m:
        tests[: n // 2] = man[rng.integers(m, size=n // 2)] + rng.normal(0, 0.03, size=(n // 2, T, 2)).astype(np.float32)
    want = DO.tube_inside(man, tests, r)
    pool = torch.from_numpy(np.concatenate([man, tests])).to(DEV)
    radius = torch.from_numpy(np.linspace(r / T, r, T)).to(DEV)
    desc = torch.tensor([(m + i, 0, m) for i in range(n)], dtype=torch.int32, device=DEV)
    got = K.tube_inside(pool, radius, desc, torch.arange(m, dtype=torch.int32, device=DEV)).cpu().numpy()
    assert np.array_equal(got, want)
    if m and n // 2:          # near-copies were planted only when n // 2 > 0
        assert 0 < want.sum()


@pytest.mark.parametrize("name", ["pred", "near"])
def test_precision_recall_cuda_matches_reference(ev, name):
    from mggan.evaluation import evaluate_precision_recall_cuda
    Kk = int(ev["meta/K"])
    got = evaluate_precision_recall_cuda(eval_ds(ev), ev[name + "/abs"], float(ev["meta/radius"]), list(range(1, Kk + 1)))
    assert len(got) == Kk + 1
    for k in got:
        assert got[k] == pytest.approx(float(ev[f"pr_{name}/{k}"]), abs=1e-12), k


# ------------------------------------------------------------------------------------------------ ADE / FDE
def test_ade_fde_cuda_matches_reference(ev):
    from mggan.evaluation import evaluate_ade_fde_cuda
    Kk = int(ev["meta/K"])
    got = evaluate_ade_fde_cuda(eval_ds(ev), ev["pred/abs"], list(range(1, Kk + 1)))
    assert len(got) == 3 * Kk
    for k in got:
        assert abs(got[k] - float(ev["ade_fde/" + k])) <= 1e-5 * max(1.0, abs(float(ev["ade_fde/" + k]))), k


def test_b200_predictions_reproduce_reference_ade_fde(ev):
    """north_star: ADE / FDE within 1e-2 of the reference from identical weights, noise and PM-Network draws --
    predictions by the B200 generator, metrics by the device kernel, against the reference's predictions + metrics."""
    import tempfile
    import mggan.model.modules.standard as S
    from mggan.evaluation import evaluate_ade_fde, evaluate_ade_fde_cuda
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    cfg = get_parser().parse_args(["--num_gens", "4", "--scene_dim", "0"])
    cfg.gpus = True
    G, D = construct_model(cfg)
    sd = {k[2:]: torch.from_numpy(v) for k, v in ev.items() if k.startswith("G/")}
    missing, unexpected = G.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("G_") for m in missing), (missing, unexpected)
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_ev_"), "ev", version=0))
    tr.G.eval()
    ds, Kk = eval_ds(ev), int(ev["meta/K"])
    idx = torch.from_numpy(ev["pred/idx"]).to(DEV)
    orig = S.MultiGenerator.get_samples
    S.MultiGenerator.get_samples = lambda self, enc_h, num_samples=5: (self.pm_logits(enc_h), idx)
    try:
        a, _, probs, gi = tr.predict(torch.from_numpy(ev["batch/in_dxdy"]).to(DEV), torch.from_numpy(ev["batch/in_xy"]).to(DEV),
                                     ds.seq_start_end, num=Kk, noise=torch.from_numpy(ev["pred/noise"]).to(DEV))
    finally:
        S.MultiGenerator.get_samples = orig
    ref = ev["pred/abs"]
    err = float((a.cpu() - torch.from_numpy(ref)).abs().max()) / float(np.abs(ref).max())
    assert err <= 1e-3, err
    nl = list(range(1, Kk + 1))
    for fn in (evaluate_ade_fde_cuda, lambda d, p, n: evaluate_ade_fde(d, p.cpu().numpy(), n)):
        got = fn(ds, a, nl)
        for k in nl:
            for m in ("ADE", "FDE"):
                assert abs(got[f"{m} k={k}"] - float(ev[f"ade_fde/{m} k={k}"])) <= 1e-2, (m, k)


def test_training_iteration_resident_images_equals_host_features():
    """The same batch once with host-cut `features` (the reference's pipeline) and once with `image_ids` + the resident
    store: identical crops, hence the same losses and parameters after one iteration."""
    import tempfile
    from collections import defaultdict
    from mggan.data_utils.scene_images import SceneImageStore
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    from mggan.synthetic import SCALING_SMALL, make_image_batch
    sizes = [3, 1, 4, 2]
    host, images = make_image_batch(sizes, seed=9)
    res, _ = make_image_batch(sizes, seed=9, resident=True)
    results = []
    for b in (host, res):
        torch.manual_seed(3)
        np.random.seed(3)
        cfg = get_parser().parse_args(["--num_gens", "2", "--num_samples", "4", "--cuda_graph", "0"])
        cfg.gpus = True
        G, D = construct_model(cfg)
        tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_ri_"), "ri", version=0))
        tr.attach_scene_images(SceneImageStore(images, SCALING_SMALL, DEV))
        tr.G.train(); tr.D.train()
        batch = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        prepared = tr._prepare(batch)
        metrics = defaultdict(list)
        torch.manual_seed(4)
        np.random.seed(4)
        tr.train_iteration(batch, metrics)
        torch.cuda.synchronize()
        results.append((prepared[5].cpu(), {k: float(v[0]) for k, v in metrics.items()},
                        # conv biases in front of train-mode BatchNorm have a zero true gradient: AdamW turns the kernels'
                        # atomic round-off into +-lr steps, so they are not compared
                        [p.detach().cpu().clone() for n, p in tr.G.named_parameters() if not n.endswith("Conv_1.bias")]))
    assert torch.equal(results[0][0], results[1][0])
    # identical inputs; the kernels' floating-point atomics make two runs agree to round-off, not bit for bit
    assert set(results[0][1]) == set(results[1][1])
    for k, v in results[0][1].items():
        assert v == pytest.approx(results[1][1][k], rel=1e-4, abs=1e-6), k
    for p, q in zip(results[0][2], results[1][2]):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-6)


def test_training_loop_deferred_crops_equal_host_features():
    """`train_iterations` stages only the image ids of a resident-image batch and the iteration cuts the crops itself --
    eagerly the first two times, then straight into the captured iteration's input on every replay.  Four iterations of
    the loop on host-cut `features` and on `image_ids` give the same losses, and the loop did replay."""
    import tempfile
    from collections import defaultdict
    from mggan.data_utils.scene_images import DeferredCrop, SceneImageStore
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    from mggan.synthetic import SCALING_SMALL, make_image_batch
    sizes = [3, 1, 4, 2]
    host, images = make_image_batch(sizes, seed=11)
    res, _ = make_image_batch(sizes, seed=11, resident=True)
    results = []
    from mggan.graph import GraphedIteration
    for b in (host, res):
        torch.manual_seed(5)
        np.random.seed(5)
        GraphedIteration._uid = 0          # the replayed sampler's Philox offsets are keyed on the capture's uid: same draws in both runs
        cfg = get_parser().parse_args(["--num_gens", "2", "--num_samples", "4", "--cuda_graph", "1"])
        cfg.gpus = True
        G, D = construct_model(cfg)
        tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_dc_"), "dc", version=0))
        tr.attach_scene_images(SceneImageStore(images, SCALING_SMALL, DEV))
        tr.G.train(); tr.D.train()
        tr.epoch = 1
        batch = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
        if "image_ids" in batch:
            assert isinstance(tr._prepare(batch, 0)[5], DeferredCrop)
        metrics = defaultdict(list)
        assert tr.train_iterations((batch for _ in range(4)), metrics) == 4
        torch.cuda.synchronize()
        assert tr.graph_hits == 2, (tr.graph_hits, tr.graph_misses)
        results.append({k: [float(x) for x in v] for k, v in metrics.items()})
    assert set(results[0]) == set(results[1])
    for k, v in results[0].items():
        assert len(v) == 4
        assert v == pytest.approx(results[1][k], rel=2e-3, abs=1e-5), k
