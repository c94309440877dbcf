"""GPU parity at BASELINE.json's full size (cfg 4: num_gens=8, k=20, 512 scenes x 32 agents = 16,384 agents) through a
size-independent property: REPLICATION INVARIANCE.  The full batch is 64 copies of an 8-scene batch (same scenes, same
scene noise, same PM-Network draws, same labels).  Every per-scene quantity of a copy then equals the one of the small
batch, train-mode BatchNorm sees the same batch statistics, and every loss of the iteration is a mean or a per-scene sum
over the global agent count -- except the two count-reweighted generator terms, which shrink by exactly 1 / copies and are
evaluated with the replicated counts on the small side -- so losses AND gradients of the 16,384-agent iteration must equal
those of the 256-agent batch, which the CPU oracle (pinned to the reference by tests/test_oracle_golden.py, premise checked
by ::test_replication_invariance_of_the_iteration) computes in seconds.  Tolerances as in tests/test_gpu_golden.py: losses 1e-3, gradients 2e-3 of the tensor's max."""
import math
from collections import defaultdict

import numpy as np
import pytest
import torch

import mggan_oracle as O
from test_gpu_golden import DEV, injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

COPIES, SCENES, AGENTS, NUM_GENS, K = 64, 8, 32, 8, 20


def check(a, b, tol, what, atol=0.0):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float((a - b).abs().max())
    bound = tol * float(b.abs().max()) + atol
    assert math.isfinite(err) and err <= bound, f"{what}: max abs err {err:.3e} > {bound:.3e}"


def grads_of(mod):
    out, seen = {}, set()
    for k, p in mod.named_parameters():
        key = k if not k.startswith("G_") else "gs." + k[2:]
        if key not in seen and p.grad is not None:
            out[key] = p.grad.detach().clone()
        seen.add(key)
    return out


def compare_grads(got, want, what):
    n = 0
    for key, v in want.items():
        if v is None or key.endswith("Conv_1.bias"):      # zero true gradient under train-mode BN: round-off on both sides
            continue
        check(got[key], v, 2e-3, f"{what} {key}", atol=1e-6)
        n += 1
    assert n >= 25, n


def test_full_size_iteration_equals_oracle_on_replicated_scenes(injected, tmp_path):  # noqa: F811
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    from mggan.synthetic import make_batch

    torch.manual_seed(17)
    cfg = get_parser().parse_args(["--num_gens", str(NUM_GENS), "--num_samples", str(K), "--cuda_graph", "0"])
    cfg.gpus = True
    G, D = construct_model(cfg)
    sdG = {k: v.detach().cpu().clone() for k, v in G.state_dict().items() if not k.startswith("G_")}
    sdD = {k: v.detach().cpu().clone() for k, v in D.state_dict().items()}
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tmp_path, "full", version=1))
    tr.epoch = 1
    tr.G.train(); tr.D.train()

    small = make_batch([AGENTS] * SCENES, seed=31, with_img=True)
    sse_s = small.pop("seq_start_end")
    small = {k: torch.from_numpy(v) for k, v in small.items()}
    n_s = SCENES * AGENTS
    big = {k: (v.repeat(COPIES, 1, 1, 1) if k == "features" else v.repeat(1, COPIES, 1)).to(DEV) for k, v in small.items()}
    sse_b = [[c * n_s + s, c * n_s + e] for c in range(COPIES) for s, e in sse_s]
    assert big["in_xy"].shape[1] == 16384 and len(sse_b) == 512

    gen = torch.Generator().manual_seed(3)
    rng = np.random.default_rng(4)

    def scene_noise():
        return torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse_s])

    d_noise, pm_noise = scene_noise(), scene_noise()
    g_noise = torch.stack([scene_noise() for _ in range(K)])
    d_idx = torch.from_numpy(rng.integers(0, NUM_GENS, size=(n_s, 1)))
    g_idx = torch.from_numpy(rng.integers(0, NUM_GENS, size=(n_s, K)))
    lab = [(float(rng.uniform(0.9, 1.0)), float(rng.uniform(0.0, 0.1))) for _ in range(3)]

    # ---- the 256-agent batch on the CPU oracle
    ob = dict(small)
    ob["seq_start_end"] = sse_s
    orc = O.OracleTrainer(sdG, sdD, NUM_GENS, num_samples=K)
    od = orc.discriminator_step(ob, d_noise[None], d_idx, lab[0], lab[1])
    # the two count-reweighted generator terms (loss / per-generator count, then mean) shrink by 1 / copies under
    # replication: the single-copy oracle is given the replicated counts (tests/test_oracle_golden.py checks this premise)
    og = orc.generator_step(ob, g_noise, g_idx, lab[2], count_scale=COPIES)
    opm = orc.net_chooser_step(ob, pm_noise[None])

    # ---- the 16,384-agent batch on the B200 path
    inj, metrics = injected, defaultdict(list)
    args = (big["in_xy"], big["in_dxdy"], big["gt_xy"], big["gt_dxdy"], sse_b, metrics, None, big["features"])
    inj.noise, inj.idx, inj.labels = [d_noise.repeat(COPIES, 1)], [d_idx.repeat(COPIES, 1)], [lab[0], lab[1]]
    tr.discriminator_step(*args)
    compare_grads(grads_of(D), od["grads"], "D step")
    check(metrics["train/info_mgan_disc_loss"][0], od["ce"], 1e-3, "ce")
    check(metrics["train/discr_loss"][0], od["real"] + od["fake"], 1e-3, "discr_loss")

    inj.noise, inj.idx, inj.labels = [g_noise.repeat(1, COPIES, 1)], [g_idx.repeat(COPIES, 1)], [lab[2]]
    tr.generator_step(*args)
    compare_grads(grads_of(G), og["grads"], "G step")
    check(metrics["train/L2_loss"][0], og["l2"], 1e-3, "l2")
    check(metrics["train/gen_loss"][0], og["adv"], 1e-3, "adv")
    check(metrics["train/info_mgan_loss"][0], og["clf"], 1e-3, "clf")

    inj.noise, inj.idx, inj.labels = [pm_noise.repeat(COPIES, 1)], [torch.zeros(n_s * COPIES, 1, dtype=torch.long)], []
    tr.net_chooser_step(*args)
    gp = grads_of(G)
    assert "gs.0.decoder.weight_hh_l0" not in gp             # the decoders run under no_grad in the PM step
    for key, v in opm["grads"].items():
        if v is not None and not key.endswith("Conv_1.bias"):
            check(gp[key], v, 2e-3, "PM step " + key, atol=1e-6)
    check(metrics["train/net_chooser_loss"][0], opm["loss"], 1e-3, "pm loss")
    assert not inj.noise and not inj.idx and not inj.labels
