"""CPU: the oracle restatement reproduces the frozen outputs of the real reference.

Fixtures were produced by `oracle/make_golden.py`, which executes the unmodified
reference modules and trainer (mggan/model/train.py:23-213,578-658).  Tolerances are
fp32 round-off only: both sides run the same PyTorch CPU kernels in (slightly)
different op order.
"""
import torch

import mggan_oracle as O


def batch_of(g):
    b = dict(g["batch"])
    b["seq_start_end"] = g["meta"]["seq_start_end"]
    return b


def close(a, b, rtol=2e-4, atol=2e-6, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e}, ref max {b.abs().max().item():.3e}"


def test_module_outputs(golden):
    g = golden
    b = batch_of(g)
    ng, k = g["meta"]["num_gens"], g["meta"]["k"]
    sdG = {n: v.clone() for n, v in g["G0"].items()}
    sdD = {n: v.clone() for n, v in g["D0"].items()}
    mask, gt_xy, gt_dxdy = O.OracleTrainer.loss_mask(b)
    img = b.get("features")
    m = g["mod"]
    with torch.no_grad():
        (rel, ab), logits, _ = O.generator_forward(sdG, ng, b["in_xy"], b["in_dxdy"], b["seq_start_end"],
                                                   m["all_noise"], True, img, 3, mask,
                                                   torch.zeros(int(mask.sum()), 3, dtype=torch.long))
        close(ab, m["all_abs"], what="all_abs")
        close(rel, m["all_rel"], what="all_rel")
        close(logits, m["logits"], what="logits")
        (rel, ab), _, _ = O.generator_forward(sdG, ng, b["in_xy"], b["in_dxdy"], b["seq_start_end"],
                                              m["sel_noise"], False, img, k, mask, m["sel_idx"])
        close(ab, m["sel_abs"], what="sel_abs")
        close(rel, m["sel_rel"], what="sel_rel")
        for mode in ("scene", "reference"):
            sd = {n: v.clone() for n, v in sdD.items()}
            o, br = O.discriminator_forward(sd, b["in_xy"], b["in_dxdy"], m["sel_abs"], m["sel_rel"],
                                            b["seq_start_end"], img, mask, social_mode=mode)
            close(o, m["d_fake_out"], what="d_fake_out " + mode)
            close(br, m["d_fake_branch"], what="d_fake_branch " + mode)
        o, br = O.discriminator_forward(sdD, b["in_xy"], b["in_dxdy"], m["sel_abs"], m["sel_rel"],
                                        b["seq_start_end"], img, mask)
        o, br = O.discriminator_forward(sdD, b["in_xy"], b["in_dxdy"], gt_xy, gt_dxdy, b["seq_start_end"], img, mask)
        close(o, m["d_real_out"], what="d_real_out")
        close(br, m["d_real_branch"], what="d_real_branch")
        (rel, ab), _, _ = O.generator_forward(sdG, ng, b["in_xy"], b["in_dxdy"], b["seq_start_end"],
                                              m["sel_noise"][:5], False, img, 5, mask, m["sel_idx"][:, :5],
                                              training=False)
        close(ab, m["eval_abs"], what="eval_abs")
    # BatchNorm buffers after the same number of train-mode forwards
    for n, v in g["Gmod"].items():
        if "running" in n or "tracked" in n:
            close(sdG[n].float(), v.float(), what="Gmod " + n)
    for n, v in g["Dmod"].items():
        if "running" in n or "tracked" in n:
            close(sdD[n].float(), v.float(), what="Dmod " + n)


def test_training_iterations(golden):
    _run_iterations(golden)


def test_training_iterations_objective_variants(golden_variant):
    """gan_obj LS / MM and weighting_target l2 / endpoint / mgan (abstract_train.py:61-79, train.py:604-647)."""
    _run_iterations(golden_variant)


def test_training_iterations_late_variants(golden_late_variant):
    """gan_type "gan": no generator-id head, no classifier terms (discriminators.py:210-211, train.py:101,181);
    pool_type "sgan": PoolHiddenNet in G and D (social_gan.py:157-229)."""
    m = golden_late_variant["meta"]
    assert (m["gan_type"] == "gan" or m["pool_type"] == "sgan" or m["experiment"] == "discrete"
            or m.get("l2_loss_type") == "mse")
    _run_iterations(golden_late_variant)


def _run_iterations(g):
    # the one decoder tensor whose optimiser state is checked: generator 0's (or the single decoder of --experiment discrete)
    dec_key = "decoder.decoder.weight_hh_l0" if g["meta"].get("experiment") == "discrete" else "gs.0.decoder.weight_hh_l0"
    b = batch_of(g)
    ng, k = g["meta"]["num_gens"], g["meta"]["k"]
    tr = O.OracleTrainer(g["G0"], g["D0"], ng, num_samples=k, gan_obj=g["meta"].get("gan_obj", "NS"),
                         weighting_target=g["meta"].get("weighting_target", "ml"),
                         l2_loss_type=g["meta"].get("l2_loss_type", "min_g_z"))
    for it in range(g["meta"]["iters"]):
        r = g[f"it{it}"]
        lab = r["labels"].tolist()
        d = tr.discriminator_step(b, r["d_noise"][None], r["d_idx"], lab[0], lab[1])
        plain = g["meta"].get("gan_type", "mgan") == "gan"
        if plain:
            assert "metric/train/info_mgan_disc_loss" not in r and "metric/train/info_mgan_loss" not in r
        else:
            close(d["ce"], r["metric/train/info_mgan_disc_loss"], what="ce")
        close(d["real"] + d["fake"], r["metric/train/discr_loss"], what="discr_loss")
        for n, v in r.items():
            if n.startswith("D_grad/"):
                close(d["grads"][n[7:]], v, rtol=2e-3, atol=1e-6, what=n)
        for u in range(1, int(g["meta"].get("num_unrolling_steps", 0)) + 1):       # abstract_train.py:139-153
            ul = r[f"labels_u{u}"].tolist()
            du = tr.discriminator_step(b, r[f"d_noise_u{u}"][None], r[f"d_idx_u{u}"], ul[0], ul[1])
            close(du["real"] + du["fake"], r[f"metric_u{u}/train/discr_loss"], what=f"discr_loss u{u}")
            for n, v in r.items():
                if n.startswith(f"D_grad_u{u}/"):
                    close(du["grads"][n.split("/", 1)[1]], v, rtol=2e-3, atol=1e-6, what=n)
        gs = tr.generator_step(b, r["g_noise"], r["g_idx"], lab[2])
        close(gs["l2"], r["metric/train/L2_loss"], what="l2")
        close(gs["adv"], r["metric/train/gen_loss"], what="adv")
        if not plain:
            close(gs["clf"], r["metric/train/info_mgan_loss"], what="clf")
        n_checked = 0
        for n, v in r.items():
            if n.startswith("G_grad/"):
                close(gs["grads"][n[7:]], v, rtol=2e-3, atol=1e-6, what=n)
                n_checked += 1
        assert n_checked >= 30
        assert gs["grads"]["net_chooser.0.weight"] is None
        pm = tr.net_chooser_step(b, r["pm_noise"][None])
        close(pm["loss"], r["metric/train/net_chooser_loss"], what="pm loss")
        for n, v in r.items():
            if n.startswith("PM_grad/"):
                close(pm["grads"][n[8:]], v, rtol=2e-3, atol=1e-6, what=n)
        assert pm["grads"][dec_key] is None
    # A conv bias that feeds a train-mode BatchNorm has an exactly-zero true gradient; what
    # reaches AdamW is round-off noise, which Adam normalises to O(lr) steps of random sign.
    # Those tensors are only required to stay within the lr envelope.
    iters = g["meta"]["iters"]
    for tag, sd in (("G1", tr.G), ("D1", tr.D)):
        for n, v in g[tag].items():
            if n.endswith("Conv_1.bias"):
                close(sd[n].detach().float(), v.float(), rtol=0, atol=2.1e-3 * 2 * iters, what=tag + " " + n)
            elif n.endswith("running_mean"):      # carries the conv bias drift x momentum
                close(sd[n].detach().float(), v.float(), rtol=1e-3, atol=5e-4 * iters, what=tag + " " + n)
            else:
                close(sd[n].detach().float(), v.float(), rtol=1e-3, atol=2e-6, what=tag + " " + n)
    st = tr.optG.state[dec_key]
    assert st["step"] == g["optG"][dec_key + "/step"] == g["meta"]["iters"]
    assert tr.optG.state["encoder.embedding.weight"]["step"] == 2 * g["meta"]["iters"]
    close(st["m"], g["optG"][dec_key + "/exp_avg"], rtol=2e-3, atol=1e-7, what="exp_avg")
    close(tr.optD.state["discs.0.0.weight"]["v"], g["optD"]["discs.0.0.weight/exp_avg_sq"], rtol=4e-3, atol=1e-10, what="exp_avg_sq")


def test_selection_indices():
    idx = torch.tensor([[1, 2, 3, 1], [0, 0, 0, 0], [2, 1, 2, 1]])
    assert O.selection_indices(idx).tolist() == [[0, 0, 0, 1], [0, 1, 2, 3], [0, 0, 1, 1]]


def test_train_then_evaluate_20_iterations():
    """ADE / FDE after TRAINING from identical weights on identical data.  20 D + G + PM iterations of the
    oracle with the reference's injected draws (tests/golden/train20.npz, oracle/make_golden_train.py), then k = 20
    predictions in eval mode: coordinates and metrics against what the unmodified reference produced."""
    from conftest import load_golden
    g = load_golden("train20")
    b = batch_of(g)
    ng, k, iters = g["meta"]["num_gens"], g["meta"]["k"], g["meta"]["iters"]
    tr = O.OracleTrainer(g["G0"], g["D0"], ng, num_samples=k)
    d = g["draws"]
    for it in range(iters):
        lab = d["labels"][it].tolist()
        tr.discriminator_step(b, d["d_noise"][it][None], d["d_idx"][it], lab[0], lab[1])
        tr.generator_step(b, d["g_noise"][it], d["g_idx"][it], lab[2])
        tr.net_chooser_step(b, d["pm_noise"][it][None])
    e = g["eval"]
    with torch.no_grad():
        (rel, ab), logits, _ = O.generator_forward(tr.G, ng, b["in_xy"], b["in_dxdy"], b["seq_start_end"], e["noise"], False,
                                                   b["features"], g["meta"]["k_eval"], None, e["idx"], training=False)
    drift = float((ab - e["abs"]).abs().max() / e["abs"].abs().max())
    assert drift <= 1e-3, drift                       # fp32 round-off through 20 Adam steps stays far below the 1e-3 bar
    m = O.ade_fde(ab, b["gt_xy"], b["seq_start_end"])
    for key in ("ADE", "FDE"):
        value, count = m[key]
        assert abs(float(value) / float(count) - float(e[key])) <= 1e-3, key


def test_replication_invariance_of_the_iteration():
    """The premise of tests/test_gpu_zb_fullsize.py, checked on the oracle itself: a batch made of C copies of the same scenes
    (same noise, same PM-Network draws, same labels) has the losses AND gradients of the single copy -- the losses are means
    (or per-scene sums over the agent count) and train-mode BatchNorm sees the same statistics -- EXCEPT the two
    count-reweighted generator terms (train.py:92-113: loss / per-generator count, then mean), which shrink by 1 / C; the
    single-copy side reproduces that with `count_scale=C`."""
    import numpy as np
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.synthetic import make_batch
    C, G, k = 3, 3, 4
    torch.manual_seed(2)
    cfg = get_parser().parse_args(["--num_gens", str(G), "--num_samples", str(k)])
    Gm, Dm = construct_model(cfg)
    sdG = {n: v.detach().clone() for n, v in Gm.state_dict().items() if not n.startswith("G_")}
    sdD = {n: v.detach().clone() for n, v in Dm.state_dict().items()}
    small = make_batch([3, 5, 2], seed=3, with_img=True)
    sse = small.pop("seq_start_end")
    small = {n: torch.from_numpy(v) for n, v in small.items()}
    n = small["in_xy"].shape[1]
    big = {n_: (v.repeat(C, 1, 1, 1) if n_ == "features" else v.repeat(1, C, 1)) for n_, v in small.items()}
    big["seq_start_end"] = [[c * n + a, c * n + b] for c in range(C) for a, b in sse]
    small["seq_start_end"] = sse
    gen, rng = torch.Generator().manual_seed(4), np.random.default_rng(5)

    def scene_noise():
        return torch.cat([torch.randn(1, 8, generator=gen).repeat(b - a, 1) for a, b in sse])

    d_noise, pm_noise, g_noise = scene_noise(), scene_noise(), torch.stack([scene_noise() for _ in range(k)])
    d_idx = torch.from_numpy(rng.integers(0, G, size=(n, 1)))
    g_idx = torch.from_numpy(rng.integers(0, G, size=(n, k)))
    lab = [(0.95, 0.04), (0.92, 0.07), (0.97, 0.02)]
    res = []
    for b, rep in ((small, 1), (big, C)):
        tr = O.OracleTrainer(sdG, sdD, G, num_samples=k)
        d = tr.discriminator_step(b, d_noise.repeat(rep, 1)[None], d_idx.repeat(rep, 1), lab[0], lab[1])
        g = tr.generator_step(b, g_noise.repeat(1, rep, 1), g_idx.repeat(rep, 1), lab[2], count_scale=C // rep)
        pm = tr.net_chooser_step(b, pm_noise.repeat(rep, 1)[None])
        res.append((d, g, pm))
    (d1, g1, p1), (d2, g2, p2) = res
    for key in ("ce", "real", "fake"):
        close(d2[key], d1[key], rtol=1e-5, atol=1e-7, what=key)
    for key in ("l2", "adv", "clf"):
        close(g2[key], g1[key], rtol=1e-5, atol=1e-7, what=key)
    close(p2["loss"], p1["loss"], rtol=1e-5, atol=1e-7, what="pm")
    for a, b, what in ((d1, d2, "D"), (g1, g2, "G"), (p1, p2, "PM")):
        for name, v in a["grads"].items():
            if v is None or name.endswith("Conv_1.bias"):
                assert (b["grads"][name] is None) == (v is None)
                continue
            close(b["grads"][name], v, rtol=2e-3, atol=2e-6, what=f"{what} {name}")
