"""GPU parity against the golden vectors frozen from the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/make_golden.py): module outputs, the gradients left after
each of the three optimisation steps, the logged loss scalars, the parameters / BatchNorm buffers
and AdamW state after whole iterations.  Random draws (scene noise, generator indices, smoothed
labels) are injected exactly as they were into the reference.

Tolerances: north-star 1e-3 relative (fp32) on predicted coordinates and discriminator outputs."""
import math
from argparse import Namespace
from collections import defaultdict

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def check(a, b, tol, what, atol=0.0):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    err = float((a - b).abs().max())
    bound = tol * float(b.abs().max()) + atol
    assert math.isfinite(err) and err <= bound, f"{what}: max abs err {err:.3e} > {bound:.3e} (ref max {float(b.abs().max()):.3e})"


def check_elementwise(a, b, what, rtol=1e-3, atol=1e-3):
    """north_star bar read element by element: |got - ref| <= atol + rtol |ref| for EVERY predicted coordinate (metres;
    atol = 1 mm), not only against the largest coordinate of the tensor."""
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    excess = (a - b).abs() - (atol + rtol * b.abs())
    worst = int(excess.argmax())
    assert float(excess.max()) <= 0.0, (f"{what}: element {worst}: got {float(a.flatten()[worst]):.6f} ref "
                                        f"{float(b.flatten()[worst]):.6f} (max abs err {float((a - b).abs().max()):.3e})")


def make_config(g):
    from mggan.model.config import get_parser
    args = get_parser().parse_args(["--num_gens", str(g["meta"]["num_gens"]), "--num_samples", str(g["meta"]["k"]),
                                    "--scene_dim", "64" if g["meta"]["with_img"] else "0",
                                    "--gan_obj", g["meta"].get("gan_obj", "NS"),
                                    "--weighting_target", g["meta"].get("weighting_target", "ml"),
                                    "--gan_type", g["meta"].get("gan_type", "mgan"),
                                    "--pool_type", g["meta"].get("pool_type", "sways"),
                                    "--experiment", g["meta"].get("experiment", "multi_generator"),
                                    "--l2_loss_type", g["meta"].get("l2_loss_type", "min_g_z"),
                                    "--num_unrolling_steps", str(g["meta"].get("num_unrolling_steps", 0))])
    args.gpus = True
    return args


def build(g):
    from mggan.model.model_factory import construct_model
    cfg = make_config(g)
    G, D = construct_model(cfg)
    G.load_state_dict(g["G0"], strict=False)
    D.load_state_dict(g["D0"], strict=False)
    return G.to(DEV).train(), D.to(DEV).train(), cfg


def batch_of(g):
    b = {k: v.to(DEV) for k, v in g["batch"].items()}
    sse = g["meta"]["seq_start_end"]
    mask = ~torch.isnan(b["gt_xy"]).any(2).any(0)
    return b, sse, mask, b["gt_xy"][:, mask], b["gt_dxdy"][:, mask]


class Injector:
    def __init__(self):
        self.noise, self.idx, self.labels = [], [], []

    def global_noise(self, dim, sub_batches, noise_type, device=None, num_samples=None):
        z = self.noise.pop(0).to(DEV)
        if num_samples is not None and z.dim() == 2:
            z = z[None]
        return z

    def gan_labels(self, shape, smoothness=0.1):
        real, fake = self.labels.pop(0)
        return torch.zeros(shape) + real, torch.zeros(shape) + fake


@pytest.fixture
def injected(monkeypatch):
    import mggan.model.modules.standard as S
    import mggan.model.train as T
    inj = Injector()
    monkeypatch.setattr(T, "get_global_noise", inj.global_noise)
    monkeypatch.setattr(S, "get_global_noise", inj.global_noise)
    monkeypatch.setattr(T, "get_gan_labels", inj.gan_labels)

    def get_samples(self, enc_h, num_samples=5):
        return self.pm_logits(enc_h), inj.idx.pop(0).to(DEV)

    monkeypatch.setattr(S.MultiGenerator, "get_samples", get_samples)
    return inj


def test_state_dict_layout_matches_reference(golden):
    """Same keys and shapes as the reference checkpoint (decoders under gs.{i} and G_{i})."""
    G, D, _ = build(golden)
    sdG, sdD = G.state_dict(), D.state_dict()
    ref_keys = set(golden["G0"])
    ours = {k for k in sdG if not k.startswith("G_")}
    assert ours == ref_keys, (sorted(ours - ref_keys), sorted(ref_keys - ours))
    for i in range(golden["meta"]["num_gens"]):
        assert f"G_{i}.decoder.weight_hh_l0" in sdG
        assert sdG[f"G_{i}.decoder.weight_hh_l0"].data_ptr() == sdG[f"gs.{i}.decoder.weight_hh_l0"].data_ptr()
    assert set(sdD) == set(golden["D0"])
    for k, v in golden["G0"].items():
        assert tuple(sdG[k].shape) == tuple(v.shape), k
    for k, v in golden["D0"].items():
        assert tuple(sdD[k].shape) == tuple(v.shape), k


def test_module_outputs(golden, injected):
    g, inj = golden, injected
    G, D, cfg = build(g)
    b, sse, mask, gt_xy, gt_dxdy = batch_of(g)
    img = b.get("features")
    m = g["mod"]
    n_act = int(mask.sum())
    k = g["meta"]["k"]
    with torch.no_grad():
        inj.idx.append(torch.zeros(n_act, 3, dtype=torch.long))
        (rel, ab), logits, _ = G(b["in_xy"], b["in_dxdy"], sse, noise=m["all_noise"].to(DEV), all_gen_out=True,
                                 img=img, num_samples=3, mask=mask)
        check(ab, m["all_abs"], 1e-3, "all_abs")
        check_elementwise(ab, m["all_abs"], "all_abs")
        check(rel, m["all_rel"], 1e-3, "all_rel")
        check_elementwise(rel, m["all_rel"], "all_rel", atol=1e-4)
        check(logits, m["logits"], 1e-3, "logits")
        inj.idx.append(m["sel_idx"])
        (rel, ab), _, _ = G(b["in_xy"], b["in_dxdy"], sse, noise=m["sel_noise"].to(DEV), all_gen_out=False, img=img,
                            num_samples=k, mask=mask)
        check(ab, m["sel_abs"], 1e-3, "sel_abs")
        check_elementwise(ab, m["sel_abs"], "sel_abs")
        check(rel, m["sel_rel"], 1e-3, "sel_rel")
        check_elementwise(rel, m["sel_rel"], "sel_rel", atol=1e-4)
        o, br = D(b["in_xy"], b["in_dxdy"], m["sel_abs"].to(DEV), m["sel_rel"].to(DEV), sse, img=img, mask=mask)
        check(o, m["d_fake_out"], 1e-3, "d_fake_out")
        check(br, m["d_fake_branch"], 1e-3, "d_fake_branch")
        o, br = D(b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, img=img, mask=mask)
        check(o, m["d_real_out"], 1e-3, "d_real_out")
        check(br, m["d_real_branch"], 1e-3, "d_real_branch")
        G.eval()
        inj.idx.append(m["sel_idx"][:, :5].contiguous())
        (rel, ab), _, _ = G(b["in_xy"], b["in_dxdy"], sse, noise=m["sel_noise"][:5].to(DEV), all_gen_out=False,
                            img=img, num_samples=5, mask=mask)
        check(ab, m["eval_abs"], 1e-3, "eval_abs")
        check_elementwise(ab, m["eval_abs"], "eval_abs")
    # BatchNorm running statistics after the same number of train-mode forwards (G: 2, D: 2 here; the
    # reference ran 3 G forwards in train mode before the eval one -> compare D only, and G's count)
    for n, v in g["Dmod"].items():
        if "running" in n or "tracked" in n:
            check(D.state_dict()[n].float(), v.float(), 1e-4, "Dmod " + n)


def test_training_iterations(golden, injected, tmp_path):
    _run_iterations(golden, injected, tmp_path)


def test_training_iterations_objective_variants(golden_variant, injected, tmp_path):
    """gan_obj LS / MM and weighting_target l2 / endpoint / mgan (abstract_train.py:61-79, train.py:604-647)."""
    _run_iterations(golden_variant, injected, tmp_path)


def _run_iterations(g, inj, tmp_path):
    # the one decoder tensor whose optimiser state is checked: generator 0's (or the single decoder of --experiment discrete)
    dec_key = "decoder.decoder.weight_hh_l0" if g["meta"].get("experiment") == "discrete" else "gs.0.decoder.weight_hh_l0"
    from mggan.logging import Experiment
    from mggan.model.train import PiNetMultiGeneratorGAN
    G, D, cfg = build(g)
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tmp_path, "golden", version=1))
    tr.epoch = 1
    b, sse, mask, gt_xy, gt_dxdy = batch_of(g)
    img = b.get("features")
    n_act = int(mask.sum())

    def grads_of(mod, max_norm=None):
        """.grad of every parameter.  max_norm: the reference's clip_grad_norm_ rescales .grad IN PLACE before the optimiser
        step (train.py:131-134, :209-212), so its frozen gradients are the clipped ones; the fused optimiser applies the
        same coefficient inside the AdamW kernel and leaves .grad untouched -- apply it here before comparing."""
        out, seen = {}, set()
        for k, p in mod.named_parameters():
            key = k if not k.startswith("G_") else "gs." + k[2:]
            if key not in seen and p.grad is not None:
                out[key] = p.grad.detach().clone()
            seen.add(key)
        if max_norm:
            total = float(torch.sqrt(sum((v.double() ** 2).sum() for v in out.values())))
            coef = min(1.0, max_norm / (total + 1e-6))
            if coef < 1.0:
                out = {k: v * coef for k, v in out.items()}
        return out

    for it in range(g["meta"]["iters"]):
        r = g[f"it{it}"]
        lab = r["labels"].tolist()
        metrics = defaultdict(list)
        inj.noise, inj.idx, inj.labels = [r["d_noise"]], [r["d_idx"]], [lab[0], lab[1]]
        tr.discriminator_step(b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        gd = grads_of(D, cfg.clipping_threshold_d)
        n = 0
        for key, v in r.items():
            if key.startswith("D_grad/"):
                if key.endswith("Conv_1.bias"):
                    continue                      # zero true gradient under train-mode BN: round-off on both sides
                check(gd[key[7:]], v, 2e-3, key, atol=1e-6)
                n += 1
        assert n >= 25
        for u in range(1, int(g["meta"].get("num_unrolling_steps", 0)) + 1):       # abstract_train.py:139-153
            ul = r[f"labels_u{u}"].tolist()
            inj.noise, inj.idx, inj.labels = [r[f"d_noise_u{u}"]], [r[f"d_idx_u{u}"]], [ul[0], ul[1]]
            tr.discriminator_step(b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
            gd = grads_of(D, cfg.clipping_threshold_d)
            for key, v in r.items():
                if key.startswith(f"D_grad_u{u}/") and not key.endswith("Conv_1.bias"):
                    check(gd[key.split("/", 1)[1]], v, 2e-3, key, atol=1e-6)
            check(metrics["train/discr_loss"][u], r[f"metric_u{u}/train/discr_loss"], 1e-3, f"discr_loss u{u}")
        plain = g["meta"].get("gan_type", "mgan") == "gan"           # no generator-id head, no classifier terms
        if plain:
            assert "train/info_mgan_disc_loss" not in metrics
        else:
            check(metrics["train/info_mgan_disc_loss"][0], r["metric/train/info_mgan_disc_loss"], 1e-3, "ce")
        check(metrics["train/discr_loss"][0], r["metric/train/discr_loss"], 1e-3, "discr_loss")

        inj.noise, inj.idx, inj.labels = [r["g_noise"]], [r["g_idx"]], [lab[2]]
        tr.generator_step(b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        gg = grads_of(G, cfg.clipping_threshold_g)
        n = 0
        for key, v in r.items():
            if key.startswith("G_grad/"):
                if key.endswith("Conv_1.bias"):
                    continue
                check(gg[key[7:]], v, 2e-3, key, atol=1e-6)
                n += 1
        assert n >= 25
        assert "net_chooser.0.weight" not in gg
        check(metrics["train/L2_loss"][0], r["metric/train/L2_loss"], 1e-3, "l2")
        check(metrics["train/gen_loss"][0], r["metric/train/gen_loss"], 1e-3, "adv")
        if plain:
            assert "train/info_mgan_loss" not in metrics
        else:
            check(metrics["train/info_mgan_loss"][0], r["metric/train/info_mgan_loss"], 1e-3, "clf")

        inj.noise, inj.idx, inj.labels = [r["pm_noise"]], [torch.zeros(n_act, 1, dtype=torch.long)], []
        tr.net_chooser_step(b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, sse, metrics, mask, img)
        gp = grads_of(G)
        for key, v in r.items():
            if key.startswith("PM_grad/"):
                if key.endswith("Conv_1.bias"):
                    continue
                check(gp[key[8:]], v, 2e-3, key, atol=1e-6)
        assert dec_key not in gp
        check(metrics["train/net_chooser_loss"][0], r["metric/train/net_chooser_loss"], 1e-3, "pm loss")
        assert not inj.noise and not inj.idx and not inj.labels

    iters = g["meta"]["iters"]
    for tag, mod in (("G1", G), ("D1", D)):
        sd = mod.state_dict()
        for n, v in g[tag].items():
            if n.endswith("Conv_1.bias"):         # Adam turns round-off noise into O(lr) steps of random sign
                check(sd[n].float(), v.float(), 0, tag + " " + n, atol=2.1e-3 * 2 * iters)
            elif n.endswith("running_mean"):
                check(sd[n].float(), v.float(), 1e-3, tag + " " + n, atol=5e-4 * iters)
            else:
                check(sd[n].float(), v.float(), 1e-3, tag + " " + n, atol=2e-6)
    p = dict(G.named_parameters())[dec_key]
    st = tr.optimizerG.state[p]
    assert float(st["step"]) == g["optG"][dec_key + "/step"] == iters
    assert float(tr.optimizerG.state[dict(G.named_parameters())["encoder.embedding.weight"]]["step"]) == 2 * iters
    check(st["exp_avg"], g["optG"][dec_key + "/exp_avg"], 2e-3, "exp_avg", atol=1e-7)
    pd_ = dict(D.named_parameters())["discs.0.0.weight"]
    check(tr.optimizerD.state[pd_]["exp_avg_sq"], g["optD"]["discs.0.0.weight/exp_avg_sq"], 4e-3, "exp_avg_sq", atol=1e-10)


def test_checkpoint_roundtrip(golden, tmp_path):
    """save() -> load_from_path() keeps the reference layout: <dir>/<name>/version_<v>/{meta_tags.csv,checkpoints}."""
    from mggan.logging import Experiment
    from mggan.model.train import PiNetMultiGeneratorGAN
    G, D, cfg = build(golden)
    w = Experiment(tmp_path, "ckpt", version=7)
    w.argparse(cfg)
    tr = PiNetMultiGeneratorGAN(G, D, cfg, w)
    tr.save(checkpoint_name="checkpoint_best.pth")
    vdir = tmp_path / "ckpt" / "version_7"
    assert (vdir / "meta_tags.csv").exists() and (vdir / "checkpoints" / "checkpoint_best.pth").exists()
    obj = torch.load(vdir / "checkpoints" / "checkpoint_best.pth", map_location="cpu")
    assert set(obj) == {"generator", "discriminator", "gen_opt", "disc_opt"}
    m2, cfg2 = PiNetMultiGeneratorGAN.load_from_path(vdir)
    assert cfg2.num_gens == cfg.num_gens
    for k, v in G.state_dict().items():
        assert torch.equal(m2.G.state_dict()[k].cpu(), v.cpu()), k
