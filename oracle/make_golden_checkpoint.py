"""TEST INFRASTRUCTURE — a checkpoint written by the UNMODIFIED reference trainer, for the drop-in boundary test
(SURVEY.md 8b "checkpoint layout must round-trip with reference checkpoints").

    python oracle/make_golden_checkpoint.py    # rewrites tests/golden/ref_checkpoint/.../checkpoint_best.pth

The reference's `PiNetMultiGeneratorGAN` (num_gens = 2, scene CNN on) runs one D + G + PM iteration on a seeded synthetic
batch so that both AdamW optimisers hold state, then its own `save()` (mggan/abstract_train.py:236-244) writes
`{"generator", "discriminator", "gen_opt", "disc_opt"}`.  `meta_tags.csv` is what test_tube's `Experiment.argparse()` +
`save()` would write next to it (key,value rows of the parsed flags; test_tube itself is not in this image, so the file
is produced here in its format, test_tube/log.py `Experiment.save`).  A few forward outputs of the saved generator are
stored beside the file so that the loader test can check the weights arrive where they belong.
"""
import os
import shutil
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
from mggan.synthetic import make_batch  # noqa: E402


def main():
    ref = refshim.load_reference()
    torch.manual_seed(404)
    np.random.seed(404)
    flags = ["--num_gens", "2", "--num_samples", "4", "--gpus", "", "--name", "refckpt"]
    args = ref.config.get_parser().parse_args(flags)
    args.gpus = False
    G, D = ref.model_factory.construct_model(args)
    tmp = tempfile.mkdtemp(prefix="mggan_ck_")
    tr = ref.train.PiNetMultiGeneratorGAN(G, D, args, ref.Experiment(tmp, "refckpt", version=3))
    tr.epoch = 1
    tr.G.train(); tr.D.train()
    b = make_batch([3, 2], seed=8, with_img=True)
    sse = b["seq_start_end"]
    t = {n: torch.from_numpy(v) for n, v in b.items() if n != "seq_start_end"}
    mask = torch.ones(t["in_xy"].shape[1], dtype=torch.bool)
    m = defaultdict(list)
    tr.discriminator_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, t["features"])
    tr.generator_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, t["features"])
    tr.net_chooser_step(t["in_xy"], t["in_dxdy"], t["gt_xy"], t["gt_dxdy"], sse, m, mask, t["features"])
    tr.save(checkpoint_name="checkpoint_best.pth")

    out_dir = os.path.join(ROOT, "tests", "golden", "ref_checkpoint", "refckpt", "version_3")
    shutil.rmtree(os.path.join(ROOT, "tests", "golden", "ref_checkpoint"), ignore_errors=True)
    os.makedirs(os.path.join(out_dir, "checkpoints"))
    src = os.path.join(tmp, "refckpt", "version_3", "checkpoints", "checkpoint_best.pth")
    shutil.copy(src, os.path.join(out_dir, "checkpoints", "checkpoint_best.pth"))
    with open(os.path.join(out_dir, "meta_tags.csv"), "w") as f:       # test_tube format: key,value
        f.write("key,value\n")
        for k, v in sorted(vars(args).items()):
            if v is None or callable(v) or isinstance(v, (list, dict)):
                continue
            if k == "gpus":
                v = True          # what the reference's main() records for a CUDA run (train.py:671-672); this run was on CPU
            f.write(f"{k},{v}\n")

    # eval-mode outputs of the saved generator on the same batch (noise and PM-Network draws injected)
    gen = torch.Generator().manual_seed(1)
    z = torch.stack([torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse]) for _ in range(4)])
    idx = torch.randint(0, 2, (t["in_xy"].shape[1], 4), generator=gen)
    orig = ref.standard.MultiGenerator.get_samples
    ref.standard.MultiGenerator.get_samples = lambda self, enc_h, num_samples=5: (orig(self, enc_h, num_samples)[0], idx)
    try:
        a, r, probs, _ = tr.predict(t["in_dxdy"], t["in_xy"], sse, img=t["features"], num=4, noise=z)
    finally:
        ref.standard.MultiGenerator.get_samples = orig
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_checkpoint", "expected.npz"),
                        in_xy=b["in_xy"], in_dxdy=b["in_dxdy"], features=b["features"],
                        seq_start_end=np.array(sse, dtype=np.int64), noise=z.numpy(), idx=idx.numpy(), abs=a.numpy(),
                        probs=probs)
    ck = torch.load(src, map_location="cpu")
    print({k: (len(v) if isinstance(v, dict) else type(v)) for k, v in ck.items()},
          os.path.getsize(src) / 1e6, "MB; opt state entries:", len(ck["gen_opt"]["state"]), len(ck["disc_opt"]["state"]))


if __name__ == "__main__":
    main()
