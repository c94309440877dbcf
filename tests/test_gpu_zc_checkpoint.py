"""GPU: `PiNetMultiGeneratorGAN.load_from_path` on a checkpoint written by the UNMODIFIED reference trainer
(tests/golden/ref_checkpoint, oracle/make_golden_checkpoint.py): the loaded generator reproduces the reference's
predictions (1e-3) and training continues from the reference's optimiser state."""
import os
import shutil
from collections import defaultdict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_checkpoint")


def test_load_reference_checkpoint_predict_and_continue(tmp_path):
    import mggan.model.modules.standard as S
    from mggan.model.train import PiNetMultiGeneratorGAN
    shutil.copytree(os.path.join(GOLD, "refckpt"), tmp_path / "refckpt")          # loading creates a writer in the folder
    tr, cfg = PiNetMultiGeneratorGAN.load_from_path(tmp_path / "refckpt" / "version_3", "best")
    assert cfg.num_gens == 2 and tr.G.n_gs == 2
    z = np.load(os.path.join(GOLD, "expected.npz"))
    sse = [tuple(int(x) for x in r) for r in z["seq_start_end"]]
    dev = {k: torch.from_numpy(z[k]).to(DEV) for k in ("in_xy", "in_dxdy", "features", "noise")}
    idx = torch.from_numpy(z["idx"]).to(DEV)
    orig = S.MultiGenerator.get_samples
    S.MultiGenerator.get_samples = lambda self, enc_h, num_samples=5: (self.pm_logits(enc_h), idx)
    try:
        a, _, probs, _ = tr.predict(dev["in_dxdy"], dev["in_xy"], sse, img=dev["features"], num=4, noise=dev["noise"])
    finally:
        S.MultiGenerator.get_samples = orig
    err = float((a.cpu() - torch.from_numpy(z["abs"])).abs().max()) / float(np.abs(z["abs"]).max())
    assert err <= 1e-3, err
    assert np.abs(probs - z["probs"]).max() <= 1e-3

    # the optimisers carry the reference's state: one more iteration advances its step counters from 1
    p = next(q for q in tr.D.parameters() if q.requires_grad)
    assert float(tr.optimizerD.state[p]["step"]) == 1.0
    from mggan.synthetic import make_batch
    b = make_batch([3, 2], seed=8, with_img=True)
    batch = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    tr.G.train(); tr.D.train()
    metrics = defaultdict(list)
    tr.train_iteration(batch, metrics)
    torch.cuda.synchronize()
    assert float(tr.optimizerD.state[p]["step"]) == 2.0
    assert all(np.isfinite(float(v[0])) for v in metrics.values())
