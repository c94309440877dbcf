"""TEST INFRASTRUCTURE — golden vectors of the reference's prediction strategies, made by EXECUTING THE
UNMODIFIED REFERENCE (`PiNetMultiGeneratorGAN.predict_expected / predict_uniform / predict_smart_sampling /
predict_rejection`, mggan/model/train.py:291-551) on seeded synthetic scenes with the scene noise injected.

    python oracle/make_golden_predict.py        # rewrites tests/golden/predict_strategies.npz

Compatibility shims, applied only while the reference runs here: `np.int = int` (train.py:310 uses the alias
numpy >= 1.24 removed).  `predict_smart_sampling` draws its generators from torch's global RNG: the draw it
returns is recorded so that the CUDA path can be given the same indices; `predict_rejection` draws its
perturbations with `torch.randn`: they are recorded by wrapping that call.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))

import refshim  # noqa: E402
from mggan.synthetic import make_batch  # noqa: E402


SIGMA = 0.7         # predict_rejection perturbation scale used for the fixture (noise perturbation = SIGMA^2 randn)


def build(ref, num_gens, with_img, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    args = ref.config.get_parser().parse_args(["--num_gens", str(num_gens), "--gpus", ""])
    args.gpus = False
    scene_dim = 64 if with_img else 0
    G = ref.standard.MultiGenerator(
        z_size=8, encoder_h_dim=32, decoder_h_dim=32, social_feat_size=32, num_gens=num_gens, pred_len=12,
        embedding_dim=16, inp_format="rel", num_social_modules=1, pool_type="sways", scene_dim=scene_dim, use_pinet=True)
    D = ref.discriminators.MultiDiscriminatorTrajectory(
        num_gens=num_gens, num_discs=1, unbound_output=False, h_dim=64, inp_format="rel", pred_len=12,
        gan_type="mgan", global_disc=1, scene_dim=scene_dim, pool_type="sways")
    with torch.no_grad():          # spread the PM-Network logits so the strategies see non-uniform probabilities
        G.net_chooser[4].weight.mul_(6.0)
        G.net_chooser[4].bias.copy_(torch.linspace(-0.8, 0.8, num_gens))
    tr = ref.train.PiNetMultiGeneratorGAN(G, D, args, ref.Experiment(tempfile.mkdtemp(prefix="mggan_gp_"), "g", version=1))
    tr.G.eval()
    return tr


def main():
    ref = refshim.load_reference()
    if not hasattr(np, "int"):
        np.int = int
    out = {}
    gen = torch.Generator().manual_seed(77)

    # ---- multi-generator strategies (G = 4, social attention only, ragged scenes incl. a single-agent one)
    tr = build(ref, 4, False, 101)
    sizes = [3, 1, 5, 2]
    b = make_batch(sizes, seed=5, with_img=False)
    sse = b["seq_start_end"]
    t = {n: torch.from_numpy(v) for n, v in b.items() if n != "seq_start_end"}
    N, Gn, num = t["in_xy"].shape[1], 4, 7
    for k_, v in tr.G.state_dict().items():
        if not k_.startswith("G_"):
            out["G4/" + k_] = v.numpy().copy()
    for n in ("in_xy", "in_dxdy"):
        out["batch4/" + n] = b[n]
    out["batch4/seq_start_end"] = np.array(sse, dtype=np.int64)

    def scene_noise(m):
        return torch.stack([torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse]) for _ in range(m)])

    z = scene_noise(num)
    a, r, p, idx = tr.predict_expected(t["in_dxdy"], t["in_xy"], sse, num=num, noise=z)
    margin = np.abs((p * num) % 1.0 - 0.5).min()
    assert margin > 1e-3, margin            # no rounding decision sits on a knife edge
    out["expected/noise"], out["expected/abs"], out["expected/rel"] = z.numpy(), a.numpy(), r.numpy()
    out["expected/probs"], out["expected/idx"] = p, idx
    zz = scene_noise(num * Gn)
    for name in ("uniform_expected", "smart_expected"):
        a, r, p, idx = tr.get_predict_func(name)(t["in_dxdy"], t["in_xy"], sse, num=num, noise=zz)
        out[name + "/noise"], out[name + "/abs"], out[name + "/rel"], out[name + "/idx"] = zz.numpy(), a.numpy(), r.numpy(), idx
    assert np.abs(p - 1.0 / Gn).min() > 1e-3
    torch.manual_seed(9)
    a, r, p, idx = tr.get_predict_func("smart_sampling")(t["in_dxdy"], t["in_xy"], sse, num=num, noise=zz)
    out["smart_sampling/noise"], out["smart_sampling/abs"], out["smart_sampling/rel"] = zz.numpy(), a.numpy(), r.numpy()
    out["smart_sampling/idx"] = idx
    assert np.abs(p - 1.0 / Gn ** 2).min() > 1e-4

    # ---- rejection (single generator, with the scene CNN in eval mode)
    tr1 = build(ref, 1, True, 202)
    b1 = make_batch([2, 3], seed=6, with_img=True)
    sse1 = b1["seq_start_end"]
    t1 = {n: torch.from_numpy(v) for n, v in b1.items() if n != "seq_start_end"}
    for k_, v in tr1.G.state_dict().items():
        if not k_.startswith("G_"):
            out["G1/" + k_] = v.numpy().copy()
    for n in ("in_xy", "in_dxdy", "features"):
        out["batch1/" + n] = b1[n]
    out["batch1/seq_start_end"] = np.array(sse1, dtype=np.int64)
    num_r = 5
    total = num_r + int(np.ceil((1 - 0.7) * num_r))
    sse_backup, sse = sse, sse1
    z1 = torch.stack([torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse1]) for _ in range(total)])
    drawn = []
    real_randn = torch.randn

    def recording_randn(*a_, **k_):
        v = real_randn(*a_, **k_)
        drawn.append(v.clone())
        return v

    torch.manual_seed(10)
    torch.randn = recording_randn
    try:
        a, r, p, idx = tr1.predict_rejection(t1["in_dxdy"], t1["in_xy"], sse1, img=t1["features"], num=num_r, noise=z1,
                                             sigma=SIGMA, N=4)
    finally:
        torch.randn = real_randn
    assert len(drawn) == 4
    # the ranking must be well conditioned (perturbations far above fp32 round-off, no near-ties), otherwise
    # parity of the kept set is not a meaningful test -- at the reference's default sigma = 1e-3 the perturbation
    # of the noise is 1e-6 and the ranking is round-off noise
    with torch.no_grad():
        base, _, _ = tr1.G(t1["in_xy"], t1["in_dxdy"], sse1, noise=z1, all_gen_out=True, img=t1["features"], num_samples=total)
        pv = base.abs.permute(3, 1, 2, 0, 4).reshape(z1.shape[1], total, -1)
        jac = torch.zeros(z1.shape[1], total)
        for e_ in drawn:
            pe, _, _ = tr1.G(t1["in_xy"], t1["in_dxdy"], sse1, noise=z1 + e_ * SIGMA ** 2, all_gen_out=True,
                             img=t1["features"], num_samples=total)
            jac += ((pe.abs.permute(3, 1, 2, 0, 4).reshape(z1.shape[1], total, -1) - pv) ** 2).sum(-1) / SIGMA ** 2
        js = torch.sort(jac / 4, dim=1).values
        gap = ((js[:, 1:] - js[:, :-1]) / js[:, 1:]).min().item()
        print("rejection: min relative gap between ranked Jacobian norms", gap)
        assert gap > 1e-2, gap
    out["rejection/noise"], out["rejection/abs"], out["rejection/rel"], out["rejection/idx"] = z1.numpy(), a.numpy(), r.numpy(), idx
    out["rejection/eps"] = torch.stack(drawn).numpy() * SIGMA ** 2         # what is added to the noise (train.py:519-521)
    out["rejection/sigma"], out["rejection/N"], out["rejection/num"] = np.float64(SIGMA), np.int64(4), np.int64(num_r)
    out["meta/num"] = np.int64(num)

    path = os.path.join(ROOT, "tests", "golden", "predict_strategies.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
