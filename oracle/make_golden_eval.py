"""TEST INFRASTRUCTURE — golden vectors of the reference's evaluation path, made
by EXECUTING THE UNMODIFIED REFERENCE in the build container:

    python oracle/make_golden_eval.py        # rewrites tests/golden/evaluation.npz and tests/golden/scene_crop.npz

evaluation.npz
  * a GoFP-shape synthetic evaluation set (every scene replicated 4 x with identical observations and different
    futures, 25 % of the agents NaN-masked), a reference `MultiGenerator` (G = 4) and the predictions of the
    reference's `PiNetMultiGeneratorGAN.predict` (mggan/model/train.py:259-289) with the scene noise and the
    PM-Network draws injected;
  * `evaluate_ade_fde` (mggan/evaluation.py:43-78) on those predictions for k = 1 .. K.  As written the function passes
    `(..., None, "raw")` positionally into `(..., mode, mode_thresh)` and raises TypeError; the
    intended call is `mode="raw"`, so `compute_metrics_from_batch` (mggan/metrics.py:99-141, unmodified) is wrapped to
    receive exactly that while the rest of `evaluate_ade_fde` runs as it is;
  * `evaluate_precision_recall` (evaluation.py:101-156) and `Manifold.compute_inside` (manifold.py:70-77) on the
    generator's predictions and on a second, better conditioned prediction set (ground-truth futures of the group +
    noise) so that both inside and outside decisions occur.

scene_crop.npz
  * `BaseDataset.ImageFeatures_small` (mggan/data_utils/BaseTrajectories.py:254-288, called unbound with the loader
    defaults margin_in = margin_out = 16, data_loaders.py:33-36,71-74,82-85) on seeded u8 RGB scene images with
    centres inside, on the border of and outside the image.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))

import refshim  # noqa: E402
from mggan.synthetic import make_batch  # noqa: E402

K = 8            # predictions per agent
RADIUS = 1.5


def make_eval(ref, ref_eval, out):
    torch.manual_seed(303)
    np.random.seed(303)
    Gn = 4
    args = ref.config.get_parser().parse_args(["--num_gens", str(Gn), "--gpus", ""])
    args.gpus = False
    G = ref.standard.MultiGenerator(
        z_size=8, encoder_h_dim=32, decoder_h_dim=32, social_feat_size=32, num_gens=Gn, pred_len=12,
        embedding_dim=16, inp_format="rel", num_social_modules=1, pool_type="sways", scene_dim=0, use_pinet=True)
    D = ref.discriminators.MultiDiscriminatorTrajectory(
        num_gens=Gn, num_discs=1, unbound_output=False, h_dim=64, inp_format="rel", pred_len=12,
        gan_type="mgan", global_disc=1, scene_dim=0, pool_type="sways")
    tr = ref.train.PiNetMultiGeneratorGAN(G, D, args, ref.Experiment(tempfile.mkdtemp(prefix="mggan_ge_"), "g", version=1))
    tr.G.eval()

    b = make_batch([3, 2, 4], seed=12, with_img=False, nan_frac=0.25, multi_future=4)
    sse = b["seq_start_end"]
    t = {n: torch.from_numpy(v) for n, v in b.items() if n != "seq_start_end"}
    N = t["in_xy"].shape[1]
    for k_, v in tr.G.state_dict().items():
        if not k_.startswith("G_"):
            out["G/" + k_] = v.numpy().copy()
    for n in ("in_xy", "in_dxdy", "gt_xy", "gt_dxdy"):
        out["batch/" + n] = b[n]
    out["batch/seq_start_end"] = np.array(sse, dtype=np.int64)

    gen = torch.Generator().manual_seed(5)
    z = torch.stack([torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse]) for _ in range(K)])
    idx = torch.randint(0, Gn, (N, K), generator=gen)
    orig = ref.standard.MultiGenerator.get_samples

    def injected(self, enc_h, num_samples=5):
        logits, _ = orig(self, enc_h, num_samples)
        return logits, idx
    ref.standard.MultiGenerator.get_samples = injected
    try:
        a, r, probs, gi = tr.predict(t["in_dxdy"], t["in_xy"], sse, num=K, noise=z)
    finally:
        ref.standard.MultiGenerator.get_samples = orig
    assert np.array_equal(gi, idx.numpy())
    preds = a.numpy()                                                   # (12, K, N, 2)
    out["pred/noise"], out["pred/idx"], out["pred/abs"], out["pred/probs"] = z.numpy(), idx.numpy(), preds, probs

    eval_ds = types.SimpleNamespace(
        pred_traj=t["gt_xy"].permute(1, 0, 2).contiguous(), obs_traj=t["in_xy"].permute(1, 0, 2).contiguous(),
        seq_start_end=[tuple(x) for x in sse], scene_list=["synthetic_gofp"] * len(sse), dataset_name="synthetic_gofp")
    n_list = list(range(1, K + 1))

    # evaluate_ade_fde with the intended mode="raw"
    real_cm = ref_eval.compute_metrics_from_batch
    ref_eval.compute_metrics_from_batch = lambda p, g, s, *a_, **k_: real_cm(p, g, s, mode="raw")
    try:
        ade = ref_eval.evaluate_ade_fde(eval_ds, preds, n_list)
    finally:
        ref_eval.compute_metrics_from_batch = real_cm
    for k_, v in ade.items():
        out["ade_fde/" + k_] = np.float64(v)

    # a second prediction set that sits around the ground-truth futures of each same-observation group
    rng = np.random.default_rng(8)
    gt = eval_ds.pred_traj.numpy()
    near = np.zeros((N, K, 12, 2), dtype=np.float32)
    groups = ref_eval.get_same_obs_indices(eval_ds)
    for same_scene in groups:
        for peds in zip(*same_scene):
            ok = [p for p in peds if not np.isnan(gt[p]).any()]
            for p in peds:
                for s in range(K):
                    src = gt[ok[rng.integers(len(ok))]] if ok else np.zeros((12, 2), np.float32)
                    scale = 0.15 if s % 2 == 0 else 1.2
                    near[p, s] = src + rng.normal(0.0, scale, size=(12, 1)).astype(np.float32) * np.linspace(0.1, 1.0, 12, dtype=np.float32)[:, None]
    near = np.ascontiguousarray(near.transpose(2, 1, 0, 3))              # (12, K, N, 2)
    out["near/abs"] = near

    for name, p in (("pred", preds), ("near", near)):
        pr = ref_eval.evaluate_precision_recall(eval_ds, p, RADIUS, n_list)
        for k_, v in pr.items():
            out[f"pr_{name}/" + k_] = np.float64(v)
        print(name, {k_: round(float(v), 4) for k_, v in pr.items()})
    vals = [out["pr_near/Precision"]] + [out[f"pr_near/Recall k={k_}"] for k_ in n_list]
    assert 0.0 < min(vals) and max(vals[:2]) < 1.0, vals                  # both decisions occur

    # raw inside masks of one manifold (all valid futures of the first group's first pedestrian)
    peds = [p for p in list(zip(*groups[0]))[0] if not np.isnan(gt[p]).any()]
    man = ref.manifold.Manifold(gt[peds], RADIUS)
    tests = near.transpose(2, 1, 0, 3)[peds].reshape(-1, 12, 2)
    inside = man.compute_inside(tests)
    out["inside/manifold"], out["inside/tests"], out["inside/mask"] = gt[peds], tests, inside
    assert 0 < inside.sum() < inside.size
    out["meta/K"], out["meta/radius"] = np.int64(K), np.float64(RADIUS)

    # knife-edge check: no distance of the "near" set sits within fp32 round-off of its radius (the device kernel
    # reproduces numpy's fp32 arithmetic exactly, this only documents that the fixture does not depend on it)
    d = np.linalg.norm(man.data[None] - tests[:, None], axis=-1)
    print("min |d - radius| :", np.abs(d - man.radius[None, None]).min())


def make_crops(out):
    from PIL import Image
    BT = refshim.load_reference_module("mggan.data_utils.BaseTrajectories")
    rng = np.random.default_rng(21)
    sizes = [(57, 41), (90, 64), (33, 33)]               # (width, height)
    scalings = [0.5, 0.7, 1.2]                           # scaling_small of the three loaders (data_loaders.py:36,74,85)
    imgs, feats, centres, img_ids, scales = [], [], [], [], []
    for i, ((w, h), sc) in enumerate(zip(sizes, scalings)):
        u8 = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        imgs.append(u8)
        pil = Image.fromarray(u8, "RGB")
        fake_self = types.SimpleNamespace(format="meter", scaling_small=sc, margin_in=16, margin_out=16)
        # centres (metres): inside, near each border, beyond the image, negative (int() truncates toward zero)
        pts = np.concatenate([
            rng.uniform([0, 0], [w * sc, h * sc], size=(6, 2)),
            np.array([[0.0, 0.0], [w * sc - 1e-3, h * sc - 1e-3], [-0.4 * sc, 3.0], [-7.3, -2.2], [w * sc + 9.0, 1.0],
                      [2.0, h * sc + 20.0], [-40.0, -40.0]])]).astype(np.float32)
        for p in pts:
            traj = torch.from_numpy(np.stack([p - 1.0, p]))           # (2 steps, 2): the last row is the centre
            f, _ = BT.BaseDataset.ImageFeatures_small(fake_self, {"small_image": pil}, traj, None)
            feats.append(f.numpy()[0])
            centres.append(p)
            img_ids.append(i)
            scales.append(sc)
    out["crop/n_images"] = np.int64(len(imgs))
    for i, u8 in enumerate(imgs):
        out[f"crop/image{i}"] = u8
    out["crop/last_xy"] = np.stack(centres).astype(np.float32)
    out["crop/image_id"] = np.array(img_ids, dtype=np.int32)
    out["crop/scaling_small"] = np.array(scales, dtype=np.float64)
    out["crop/features"] = np.stack(feats).astype(np.float32)
    assert out["crop/features"].shape[1:] == (4, 33, 33), out["crop/features"].shape


def main():
    ref = refshim.load_reference()
    ref_eval = refshim.load_reference_module("mggan.evaluation")
    gold = os.path.join(ROOT, "tests", "golden")
    out = {}
    make_eval(ref, ref_eval, out)
    np.savez_compressed(os.path.join(gold, "evaluation.npz"), **out)
    print("evaluation.npz", len(out), "arrays")
    out = {}
    make_crops(out)
    np.savez_compressed(os.path.join(gold, "scene_crop.npz"), **out)
    print("scene_crop.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
