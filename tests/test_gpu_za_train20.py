"""GPU: "train, then evaluate" parity (SURVEY.md 8d / north_star: ADE / FDE within 1e-2 after training from identical weights
on identical data).  20 D + G + PM iterations of the B200 trainer with the draws the reference used
(tests/golden/train20.npz, made by oracle/make_golden_train.py from the UNMODIFIED reference), then k = 20 predictions in
eval mode: metrics within 1e-2, coordinates within 2e-3 of the reference's.  (The CPU oracle run through the same fixture
drifts 4e-5 in the coordinates: tests/test_oracle_golden.py::test_train_then_evaluate_20_iterations.)"""
from collections import defaultdict

import pytest
import torch

from conftest import load_golden
from test_gpu_golden import DEV, batch_of, build, injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_train_20_iterations_then_ade_fde(injected, tmp_path):  # noqa: F811
    from mggan.logging import Experiment
    from mggan.metrics import compute_metrics_from_batch
    from mggan.model.train import PiNetMultiGeneratorGAN
    g = load_golden("train20")
    g["meta"]["with_img"] = 1
    G, D, cfg = build(g)
    cfg.cuda_graph = 0
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tmp_path, "train20", version=1))
    tr.epoch = 1
    b, sse, mask, gt_xy, gt_dxdy = batch_of(g)
    img = b["features"]
    n = b["in_xy"].shape[1]
    inj, d = injected, g["draws"]
    for it in range(g["meta"]["iters"]):
        lab = d["labels"][it].tolist()
        metrics = defaultdict(list)
        args = (b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        inj.noise, inj.idx, inj.labels = [d["d_noise"][it]], [d["d_idx"][it]], [lab[0], lab[1]]
        tr.discriminator_step(*args)
        inj.noise, inj.idx, inj.labels = [d["g_noise"][it]], [d["g_idx"][it]], [lab[2]]
        tr.generator_step(*args)
        inj.noise, inj.idx, inj.labels = [d["pm_noise"][it]], [torch.zeros(n, 1, dtype=torch.long)], []
        tr.net_chooser_step(*args)
        assert not inj.noise and not inj.idx and not inj.labels
    e = g["eval"]
    inj.idx = [e["idx"]]
    a, _, _, _ = tr.predict(b["in_dxdy"], b["in_xy"], sse, img=img, num=g["meta"]["k_eval"], noise=e["noise"].to(DEV))
    drift = float((a.cpu() - e["abs"]).abs().max() / e["abs"].abs().max())
    assert drift <= 2e-3, drift
    m = compute_metrics_from_batch(a.cpu(), g["batch"]["gt_xy"], sse, mode="raw")
    for key in ("ADE", "FDE"):
        value, count = m[key]
        assert abs(value / count - float(e[key])) <= 1e-2, (key, value / count, float(e[key]))
