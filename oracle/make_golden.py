"""TEST INFRASTRUCTURE — freezes golden vectors by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

For every case it builds the reference's own `MultiGenerator` /
`MultiDiscriminatorTrajectory` (mggan/model/modules/standard.py:17,
discriminators.py:12) and `PiNetMultiGeneratorGAN` trainer (mggan/model/train.py:18),
feeds a seeded synthetic batch, injects every random draw (scene noise, PM-Net
generator indices, smoothed GAN labels) and records module outputs, the gradients
left in `.grad` after each of the three steps, the loss scalars the reference logs,
and the final parameters / BatchNorm buffers.  The fixtures are what pins
`oracle/mggan_oracle.py` (tests/test_oracle_golden.py) and, on the GPU, the CUDA path.
"""
import argparse
import os
import sys
import tempfile
from collections import defaultdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))

import refshim  # noqa: E402
from mggan.synthetic import make_batch  # noqa: E402  (numpy only; no CUDA needed)

CASES = {
    # name: dict(num_gens, sizes, with_img, nan_frac, k, iters, seed)
    "cfg1_g1_tiny": dict(num_gens=1, sizes=[4], with_img=True, nan_frac=0.0, k=20, iters=1, seed=11),
    "cfg2_g4_eth_noimg": dict(num_gens=4, sizes=[1, 3, 2, 5, 1, 4], with_img=False, nan_frac=0.0, k=20, iters=1, seed=22),
    "cfg3_g8_sdd_masked": dict(num_gens=8, sizes=[4, 6, 3], with_img=True, nan_frac=0.25, k=20, iters=2, seed=33),
    # flag-reachable objective variants (abstract_train.py:61-79, train.py:604-647): iteration vectors only
    "var_ls_l2": dict(num_gens=4, sizes=[2, 3, 1], with_img=False, nan_frac=0.0, k=6, iters=1, seed=44,
                      flags=["--gan_obj", "LS", "--weighting_target", "l2"], modules=False),
    "var_mm_endpoint": dict(num_gens=4, sizes=[3, 2], with_img=False, nan_frac=0.0, k=6, iters=1, seed=55,
                            flags=["--gan_obj", "MM", "--weighting_target", "endpoint"], modules=False),
    "var_ns_mgan": dict(num_gens=3, sizes=[2, 2], with_img=True, nan_frac=0.0, k=5, iters=1, seed=66,
                        flags=["--weighting_target", "mgan"], modules=False),
    # --pool_type sgan: PoolHiddenNet instead of the Social-Ways attention in G and D (social_gan.py:157-229)
    "var_sgan_pool": dict(num_gens=3, sizes=[3, 1, 4, 2], with_img=False, nan_frac=0.0, k=5, iters=1, seed=88,
                          flags=["--pool_type", "sgan"], modules=False),
    # --experiment discrete: DiscreteLatentGenerator, one decoder + a discrete latent code (standard_discrete.py:18-257);
    # built by the reference's own construct_model (scene CNN on: model_factory.py:19 hard-codes scene_dim)
    "var_discrete": dict(num_gens=3, sizes=[2, 3, 1], with_img=True, nan_frac=0.0, k=5, iters=1, seed=99,
                         flags=["--experiment", "discrete"], modules=False),
    # l2_loss_type "mse" (train.py:58-65: squared per-step distances) + num_unrolling_steps 1 (abstract_train.py:139-165: two
    # discriminator steps per iteration; `backup = self.D.state_dict()` aliases the live tensors, so the restore after the
    # PM step is a no-op -- executed here exactly as the reference loop does)
    "var_mse_unroll": dict(num_gens=3, sizes=[2, 3, 1], with_img=True, nan_frac=0.0, k=5, iters=1, seed=111,
                           flags=["--l2_loss_type", "mse", "--num_unrolling_steps", "1"], modules=False),
    # gan_type "gan": plain discriminator, no generator-id head (discriminators.py:210-211, train.py:101,181)
    "var_gan_plain": dict(num_gens=3, sizes=[3, 1, 2], with_img=True, nan_frac=0.0, k=5, iters=1, seed=77,
                          flags=["--gan_type", "gan"], modules=False),
}


class Injector:
    """Queues that replace the reference's RNG draws, in call order."""

    def __init__(self):
        self.noise, self.idx, self.labels = [], [], []

    def global_noise(self, dim, sub_batches, noise_type):
        return self.noise.pop(0)

    def gan_labels(self, shape, smoothness=0.1):
        real, fake = self.labels.pop(0)
        return torch.zeros(shape) + real, torch.zeros(shape) + fake


def build(ref, case):
    torch.manual_seed(case["seed"])
    np.random.seed(case["seed"])
    args = ref.config.get_parser().parse_args(["--num_gens", str(case["num_gens"]), "--gpus", "",
                                               "--num_samples", str(case["k"])] + case.get("flags", []))
    args.gpus = False
    scene_dim = 64 if case["with_img"] else 0
    args.use_pinet = True
    G = ref.standard.MultiGenerator(
        z_size=8, encoder_h_dim=32, decoder_h_dim=32, social_feat_size=32, num_gens=case["num_gens"],
        pred_len=12, embedding_dim=16, inp_format="rel", num_social_modules=1, pool_type=args.pool_type,
        scene_dim=scene_dim, use_pinet=True)
    D = ref.discriminators.MultiDiscriminatorTrajectory(
        num_gens=case["num_gens"], num_discs=1, unbound_output=args.gan_obj in ["W", "LS"], h_dim=64, inp_format="rel",
        pred_len=12, gan_type=args.gan_type, global_disc=1, scene_dim=scene_dim, pool_type=args.pool_type)
    if args.experiment == "discrete":
        G, D = ref.model_factory.construct_model(args)
    elif case["with_img"]:
        # the CLI path must build the very same thing (model_factory.py:7-86)
        G2, D2 = ref.model_factory.construct_model(args)
        assert [k for k in G2.state_dict()] == [k for k in G.state_dict()]
        assert [k for k in D2.state_dict()] == [k for k in D.state_dict()]
    # make BatchNorm affine / biases non-trivial so parity exercises them
    with torch.no_grad():
        for n, p in list(G.named_parameters()) + list(D.named_parameters()):
            if "BN_1.weight" in n:
                p.copy_(torch.empty_like(p).uniform_(0.5, 1.5) * torch.where(torch.rand_like(p) < 0.25, -1.0, 1.0))
            elif "BN_1.bias" in n:
                p.copy_(torch.empty_like(p).uniform_(-0.3, 0.3))
    tmp = tempfile.mkdtemp(prefix="mggan_golden_")
    trainer = ref.train.PiNetMultiGeneratorGAN(G, D, args, ref.Experiment(tmp, "golden", version=1))
    trainer.epoch = 1
    trainer.G.train()
    trainer.D.train()
    return trainer, args


def sd_np(prefix, module, out):
    for k, v in module.state_dict().items():
        if k.startswith("G_"):
            continue
        out[f"{prefix}/{k}"] = v.detach().cpu().numpy().copy()


def grads_np(prefix, module, out):
    seen = set()
    for k, p in module.named_parameters():
        key = k if not k.startswith("G_") else "gs." + k[2:]
        if key in seen:
            continue
        seen.add(key)
        if p.grad is not None:
            out[f"{prefix}/{key}"] = p.grad.detach().cpu().numpy().copy()


def run_case(ref, name, case):
    trainer, args = build(ref, case)
    G, D = trainer.G, trainer.D
    k, ng = case["k"], case["num_gens"]
    b = make_batch(case["sizes"], seed=case["seed"], with_img=case["with_img"], nan_frac=case["nan_frac"])
    sse = b["seq_start_end"]
    t = {n: torch.from_numpy(v) for n, v in b.items() if n != "seq_start_end"}
    img = t.get("features")
    N = t["in_xy"].shape[1]
    mask = ~t["gt_xy"].isnan().any(2).any(0)
    n_act = int(mask.sum())
    gt_xy, gt_dxdy = t["gt_xy"][:, mask], t["gt_dxdy"][:, mask]

    out = {"meta/gan_obj": np.array(args.gan_obj), "meta/weighting_target": np.array(args.weighting_target),
           "meta/gan_type": np.array(args.gan_type), "meta/pool_type": np.array(args.pool_type),
           "meta/experiment": np.array(args.experiment), "meta/l2_loss_type": np.array(args.l2_loss_type),
           "meta/num_unrolling_steps": np.int64(args.num_unrolling_steps),
           "meta/num_gens": np.int64(ng), "meta/k": np.int64(k), "meta/iters": np.int64(case["iters"]),
           "meta/with_img": np.int64(case["with_img"]), "meta/seq_start_end": np.array(sse, dtype=np.int64)}
    for n, v in b.items():
        if n != "seq_start_end":
            out[f"batch/{n}"] = v
    sd_np("G0", G, out)
    sd_np("D0", D, out)

    inj = Injector()
    ref.train.get_global_noise = inj.global_noise
    ref.standard.get_global_noise = inj.global_noise
    ref.train.get_gan_labels = inj.gan_labels
    forced = {}

    def get_samples(self, enc_h, num_samples=5):
        logits = self.net_chooser(enc_h) if self.use_pinet else self.net_prior.expand(enc_h.size(0), -1)
        return logits, inj.idx.pop(0)

    ref.standard.MultiGenerator.get_samples = get_samples
    # the discrete-latent generator lives in a module the shim does not expose: patch it through its globals
    DLG = ref.model_factory.DiscreteLatentGenerator
    DLG.forward.__globals__["get_global_noise"] = inj.global_noise
    DLG.get_samples = get_samples
    rng = np.random.default_rng(case["seed"] + 1)
    gen = torch.Generator().manual_seed(case["seed"] + 2)

    def scene_noise():
        return torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse])

    if case.get("modules", True):
        # ---- module-level vectors on the initial weights (BN buffers restored afterwards)
        g_state = {k_: v.clone() for k_, v in G.state_dict().items()}
        d_state = {k_: v.clone() for k_, v in D.state_dict().items()}
        with torch.no_grad():
            z3 = torch.stack([scene_noise() for _ in range(3)])
            inj.idx.append(torch.zeros(n_act, 3, dtype=torch.long))
            (rel, ab), logits, _ = G(t["in_xy"], t["in_dxdy"], sse, noise=z3, all_gen_out=True, img=img,
                                     num_samples=3, mask=mask)
            out["mod/all_noise"], out["mod/all_abs"], out["mod/all_rel"] = z3.numpy(), ab.numpy(), rel.numpy()
            out["mod/logits"] = logits.numpy()
            zk = torch.stack([scene_noise() for _ in range(k)])
            idx = torch.from_numpy(rng.integers(0, ng, size=(n_act, k)))
            inj.idx.append(idx)
            (rel, ab), _, _ = G(t["in_xy"], t["in_dxdy"], sse, noise=zk, all_gen_out=False, img=img,
                                num_samples=k, mask=mask)
            out["mod/sel_noise"], out["mod/sel_idx"] = zk.numpy(), idx.numpy()
            out["mod/sel_abs"], out["mod/sel_rel"] = ab.numpy(), rel.numpy()
            o, br = D(t["in_xy"], t["in_dxdy"], ab, rel, sse, img=img, mask=mask)
            out["mod/d_fake_out"], out["mod/d_fake_branch"] = o.numpy(), br.numpy()
            o, br = D(t["in_xy"], t["in_dxdy"], gt_xy, gt_dxdy, sse, img=img, mask=mask)
            out["mod/d_real_out"], out["mod/d_real_branch"] = o.numpy(), br.numpy()
            # eval-mode generator (BatchNorm running statistics), as used by predict() train.py:259-289
            G.eval()
            inj.idx.append(idx[:, :5])
            (rel, ab), _, _ = G(t["in_xy"], t["in_dxdy"], sse, noise=zk[:5], all_gen_out=False, img=img,
                                num_samples=5, mask=mask)
            out["mod/eval_abs"] = ab.numpy()
            G.train()
        sd_np("Gmod", G, out)       # BN buffers after 3 train-mode forwards of G-CNN / 2 of D-CNN
        sd_np("Dmod", D, out)
        G.load_state_dict(g_state)
        D.load_state_dict(d_state)

    # ---- full iterations: D step, G step, PM step (abstract_train.py:136-159)
    for it in range(case["iters"]):
        P = f"it{it}"
        metrics = defaultdict(list)
        d_noise = scene_noise()
        d_idx = torch.from_numpy(rng.integers(0, ng, size=(n_act, 1)))
        lab = [(float(rng.uniform(0.9, 1.0)), float(rng.uniform(0.0, 0.1))) for _ in range(3)]
        g_noise = torch.stack([scene_noise() for _ in range(k)])
        g_idx = torch.from_numpy(rng.integers(0, ng, size=(n_act, k)))
        pm_noise = scene_noise()
        out[f"{P}/d_noise"], out[f"{P}/d_idx"] = d_noise.numpy(), d_idx.numpy()
        out[f"{P}/labels"] = np.array(lab, dtype=np.float64)          # rows: D-real pass, D-fake pass, G step; cols (real, fake)
        out[f"{P}/g_noise"], out[f"{P}/g_idx"], out[f"{P}/pm_noise"] = g_noise.numpy(), g_idx.numpy(), pm_noise.numpy()

        inj.noise, inj.idx, inj.labels = [d_noise.clone()], [d_idx], [lab[0], lab[1]]
        trainer.discriminator_step(t["in_xy"], t["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        grads_np(f"{P}/D_grad", D, out)
        backup = None
        for u in range(1, args.num_unrolling_steps + 1):          # abstract_train.py:139-153
            if u == 1:
                backup = trainer.D.state_dict()
            un, ui = scene_noise(), torch.from_numpy(rng.integers(0, ng, size=(n_act, 1)))
            ul = [(float(rng.uniform(0.9, 1.0)), float(rng.uniform(0.0, 0.1))) for _ in range(2)]
            out[f"{P}/d_noise_u{u}"], out[f"{P}/d_idx_u{u}"] = un.numpy(), ui.numpy()
            out[f"{P}/labels_u{u}"] = np.array(ul, dtype=np.float64)
            inj.noise, inj.idx, inj.labels = [un.clone()], [ui], [ul[0], ul[1]]
            trainer.discriminator_step(t["in_xy"], t["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
            grads_np(f"{P}/D_grad_u{u}", D, out)

        inj.noise, inj.idx, inj.labels = [z.clone() for z in g_noise], [g_idx], [lab[2]]
        trainer.generator_step(t["in_xy"], t["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        grads_np(f"{P}/G_grad", G, out)

        inj.noise, inj.idx, inj.labels = [pm_noise.clone()], [torch.zeros(n_act, 1, dtype=torch.long)], []
        trainer.net_chooser_step(t["in_xy"], t["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        grads_np(f"{P}/PM_grad", G, out)
        if backup is not None:
            trainer.D.load_state_dict(backup)                         # abstract_train.py:161-162
        for mk, mv in metrics.items():
            if not mk.startswith("probs/"):
                out[f"{P}/metric/{mk}"] = np.float64(mv[0])
                for u in range(1, len(mv)):                           # the unrolled discriminator steps log again
                    out[f"{P}/metric_u{u}/{mk}"] = np.float64(mv[u])
        assert not inj.noise and not inj.idx and not inj.labels

    sd_np("G1", G, out)
    sd_np("D1", D, out)
    # AdamW state of one decoder tensor and one encoder tensor (step counters differ: SURVEY App. B)
    dec_key = "gs.0.decoder.weight_hh_l0" if args.experiment != "discrete" else "decoder.decoder.weight_hh_l0"
    for tag, opt, mod, key in (("G", trainer.optimizerG, G, "encoder.embedding.weight"),
                               ("G", trainer.optimizerG, G, dec_key),
                               ("D", trainer.optimizerD, D, "discs.0.0.weight")):
        p = dict(mod.named_parameters())[key]
        st = opt.state[p]
        kk = key if not key.startswith("G_") else "gs." + key[2:]
        out[f"opt{tag}/{kk}/step"] = np.float64(float(st["step"]))
        out[f"opt{tag}/{kk}/exp_avg"] = st["exp_avg"].numpy().copy()
        out[f"opt{tag}/{kk}/exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()

    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: N={N} N_act={n_act} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.set_num_threads(8)
    ref = refshim.load_reference()
    for name, case in CASES.items():
        if a.only and a.only != name:
            continue
        run_case(ref, name, case)
