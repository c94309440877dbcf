"""GPU parity of the prediction strategies (`PiNetMultiGeneratorGAN.get_predict_func`, reference
mggan/model/train.py:291-576) against vectors frozen from the UNMODIFIED reference
(tests/golden/predict_strategies.npz, oracle/make_golden_predict.py): same weights, same scene noise; the
generator indices chosen by each strategy must be identical and the predicted coordinates within 1e-3."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "predict_strategies.npz")


def check(a, b, tol, what):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float((a - b).abs().max())
    assert err <= tol * float(b.abs().max()), f"{what}: {err:.3e} vs max {float(b.abs().max()):.3e}"


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


def trainer(gold, prefix, num_gens, with_img):
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    cfg = get_parser().parse_args(["--num_gens", str(num_gens), "--scene_dim", "64" if with_img else "0"])
    cfg.gpus = True
    G, D = construct_model(cfg)
    sd = {k[len(prefix) + 1:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith(prefix + "/")}
    missing, unexpected = G.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("G_") for m in missing), (missing, unexpected)
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_ps_"), "ps", version=0))
    tr.G.eval()
    return tr


def batch(gold, prefix):
    b = {k.split("/")[1]: torch.from_numpy(v).to(DEV) for k, v in gold.items()
         if k.startswith(prefix + "/") and "seq_start_end" not in k}
    return b, [tuple(int(x) for x in r) for r in gold[prefix + "/seq_start_end"]]


@pytest.mark.parametrize("name", ["expected", "uniform_expected", "smart_expected"])
def test_deterministic_strategies_match_reference(gold, name):
    tr = trainer(gold, "G4", 4, False)
    b, sse = batch(gold, "batch4")
    num = int(gold["meta/num"])
    noise = torch.from_numpy(gold[name + "/noise"]).to(DEV)
    a, r, probs, idx = tr.get_predict_func(name)(b["in_dxdy"], b["in_xy"], sse, num=num, noise=noise)
    assert np.array_equal(idx, gold[name + "/idx"]), (idx, gold[name + "/idx"])
    check(a, gold[name + "/abs"], 1e-3, name + " abs")
    check(r, gold[name + "/rel"], 1e-3, name + " rel")
    if name == "expected":
        check(probs, gold["expected/probs"], 1e-4, "probs")


def test_smart_sampling_matches_reference_given_its_draw(gold, monkeypatch):
    import mggan.model.train as T
    tr = trainer(gold, "G4", 4, False)
    b, sse = batch(gold, "batch4")
    num = int(gold["meta/num"])
    drawn = torch.from_numpy(gold["smart_sampling/idx"]).to(DEV)
    allowed = {}

    def fake(probs, num, eps):
        allowed["over"] = (probs > eps).cpu()
        return drawn

    monkeypatch.setattr(T, "threshold_sample_indices", fake)
    noise = torch.from_numpy(gold["smart_sampling/noise"]).to(DEV)
    a, r, _, idx = tr.get_predict_func("smart_sampling")(b["in_dxdy"], b["in_xy"], sse, num=num, noise=noise)
    assert np.array_equal(idx, gold["smart_sampling/idx"])
    # the reference's draw only used generators over the threshold our probabilities give
    assert bool(allowed["over"].gather(1, drawn.cpu()).all())
    check(a, gold["smart_sampling/abs"], 1e-3, "smart_sampling abs")
    check(r, gold["smart_sampling/rel"], 1e-3, "smart_sampling rel")
    # and the un-patched sampler runs on the device
    monkeypatch.undo()
    a2, _, _, idx2 = tr.get_predict_func("uniform_sampling")(b["in_dxdy"], b["in_xy"], sse, num=num)
    assert a2.shape == a.shape and idx2.shape == idx.shape and 0 <= idx2.min() and idx2.max() < 4


def test_rejection_matches_reference(gold):
    tr = trainer(gold, "G1", 1, True)
    b, sse = batch(gold, "batch1")
    noise = torch.from_numpy(gold["rejection/noise"]).to(DEV)
    eps = torch.from_numpy(gold["rejection/eps"]).to(DEV)
    a, r, _, idx = tr.predict_rejection(b["in_dxdy"], b["in_xy"], sse, img=b["features"], num=int(gold["rejection/num"]),
                                        noise=noise, sigma=float(gold["rejection/sigma"]), N=int(gold["rejection/N"]),
                                        eps_noise=eps)
    assert np.array_equal(idx, gold["rejection/idx"])
    check(a, gold["rejection/abs"], 1e-3, "rejection abs")
    check(r, gold["rejection/rel"], 1e-3, "rejection rel")
