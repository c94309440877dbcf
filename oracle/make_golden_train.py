"""TEST INFRASTRUCTURE — "train, then evaluate" parity fixture (ADE / FDE @ k = 20 within 1e-2 between the
reference and the B200 path TRAINED from identical weights on identical data), made by EXECUTING THE UNMODIFIED REFERENCE:

    python oracle/make_golden_train.py        # rewrites tests/golden/train20.npz

The reference's PiNetMultiGeneratorGAN (G = 3, scene CNN on) runs 20 D + G + PM iterations on a seeded synthetic batch with
every random draw injected (scene noise, PM-Network draws, smoothed labels), then predicts k = 20 futures in eval mode
(noise and draws injected) and scores them with its own `compute_metrics_from_batch` (mggan/metrics.py:99-141).  Stored:
initial weights, the batch, every injected draw, the final predictions and the metrics.
"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))

import refshim  # noqa: E402
import make_golden as MG  # noqa: E402
from mggan.synthetic import make_batch  # noqa: E402

CASE = dict(num_gens=3, sizes=[3, 2, 4], with_img=True, nan_frac=0.0, k=5, iters=20, seed=123)
K_EVAL = 20


def main():
    ref = refshim.load_reference()
    trainer, args = MG.build(ref, CASE)
    G, D = trainer.G, trainer.D
    k, ng = CASE["k"], CASE["num_gens"]
    b = make_batch(CASE["sizes"], seed=CASE["seed"], with_img=True)
    sse = b["seq_start_end"]
    t = {n: torch.from_numpy(v) for n, v in b.items() if n != "seq_start_end"}
    img = t["features"]
    N = t["in_xy"].shape[1]
    mask = torch.ones(N, dtype=torch.bool)
    out = {"meta/num_gens": np.int64(ng), "meta/k": np.int64(k), "meta/iters": np.int64(CASE["iters"]),
           "meta/k_eval": np.int64(K_EVAL), "meta/seq_start_end": np.array(sse, dtype=np.int64)}
    for n, v in b.items():
        if n != "seq_start_end":
            out[f"batch/{n}"] = v
    MG.sd_np("G0", G, out)
    MG.sd_np("D0", D, out)

    inj = MG.Injector()
    ref.train.get_global_noise = inj.global_noise
    ref.standard.get_global_noise = inj.global_noise
    ref.train.get_gan_labels = inj.gan_labels
    ref.standard.MultiGenerator.get_samples = lambda self, enc_h, num_samples=5: (self.net_chooser(enc_h), inj.idx.pop(0))
    rng = np.random.default_rng(CASE["seed"] + 1)
    gen = torch.Generator().manual_seed(CASE["seed"] + 2)

    def scene_noise():
        return torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse])

    draws = defaultdict(list)
    for it in range(CASE["iters"]):
        d_noise, pm_noise = scene_noise(), scene_noise()
        g_noise = torch.stack([scene_noise() for _ in range(k)])
        d_idx = torch.from_numpy(rng.integers(0, ng, size=(N, 1)))
        g_idx = torch.from_numpy(rng.integers(0, ng, size=(N, k)))
        lab = [(float(rng.uniform(0.9, 1.0)), float(rng.uniform(0.0, 0.1))) for _ in range(3)]
        for key, v in (("d_noise", d_noise), ("pm_noise", pm_noise), ("g_noise", g_noise), ("d_idx", d_idx), ("g_idx", g_idx),
                       ("labels", torch.tensor(lab, dtype=torch.float64))):
            draws[key].append(v.numpy())
        m = defaultdict(list)
        inj.noise, inj.idx, inj.labels = [d_noise.clone()], [d_idx], [lab[0], lab[1]]
        trainer.discriminator_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, img)
        inj.noise, inj.idx, inj.labels = [z.clone() for z in g_noise], [g_idx], [lab[2]]
        trainer.generator_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, img)
        inj.noise, inj.idx, inj.labels = [pm_noise.clone()], [torch.zeros(N, 1, dtype=torch.long)], []
        trainer.net_chooser_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, img)
        assert not inj.noise and not inj.idx and not inj.labels
        if it in (0, CASE["iters"] - 1):
            print(it, {kk: round(float(v[0]), 4) for kk, v in m.items() if kk.startswith("train/")})
    for key, v in draws.items():
        out["draws/" + key] = np.stack(v)

    z = torch.stack([scene_noise() for _ in range(K_EVAL)])
    e_idx = torch.from_numpy(rng.integers(0, ng, size=(N, K_EVAL)))
    inj.idx = [e_idx]
    a, r, probs, _ = trainer.predict(t["in_dxdy"], t["in_xy"], sse, img=img, num=K_EVAL, noise=z)
    met = ref.metrics.compute_metrics_from_batch(a, t["gt_xy"], sse, mode="raw")
    out["eval/noise"], out["eval/idx"], out["eval/abs"], out["eval/probs"] = z.numpy(), e_idx.numpy(), a.numpy(), probs
    for key, (value, count) in met.items():
        out[f"eval/{key}"] = np.float64(value / count)
    print({key: float(out[f"eval/{key}"]) for key in met})
    MG.sd_np("G1", G, out)
    path = os.path.join(ROOT, "tests", "golden", "train20.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
