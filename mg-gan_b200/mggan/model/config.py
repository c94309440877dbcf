"""Command-line flags of `mggan.model.train` (reference: mggan/model/config.py:4-135).

The flag set is a drop-in contract: every reference flag is kept with its name, type, default
and choices (checkpoints store them in meta_tags.csv and `load()` merges them over these
defaults).  The reference builds them on test_tube's HyperOptArgumentParser; only `opt_list`
(argparse + a list of tunable options) is used from it, so a plain argparse subclass suffices.
Flags added by this implementation are additive and listed at the end.
"""
import argparse


class HyperOptArgumentParser(argparse.ArgumentParser):
    """argparse with test_tube's `opt_list` (options / tunable are accepted and ignored)."""

    def __init__(self, strategy="grid_search", **kw):
        super().__init__(**kw)
        self.strategy = strategy

    def opt_list(self, *args, options=None, tunable=False, **kw):
        return self.add_argument(*args, **kw)


DATASETS = ["hotel", "eth", "zara1", "zara2", "univ", "social_stanford_synthetic", "stanford", "gofp"]

# (flag, kwargs) in the reference's order
_FLAGS = [
    ("--name", dict(type=str, default="test")),
    ("--log_dir", dict(type=str, default="./logs/")),
    # default "stanford_synthetic" is not among its own choices in the reference either (argparse
    # does not validate defaults); the synthetic_* names are this implementation's generators
    ("--dataset", dict(type=str, default="stanford_synthetic",
                       choices=DATASETS + ["synthetic_tiny", "synthetic_eth", "synthetic_sdd", "synthetic_univ",
                                           "synthetic_gofp"])),
    ("--gpus", dict(type=str, default="0")),
    ("--workers", dict(type=int, default=0)),
    ("--batch_size", dict(type=int, default=2)),
    ("--beta1", dict(type=float, default=0.5, opt=[0.1, 0.5, 0.9])),
    ("--l2_loss_weight", dict(type=float, default=1.0)),
    ("--clf_loss_weight", dict(type=float, default=1.0)),
    ("--pi_net_loss_weight", dict(type=float, default=1.0)),
    ("--epochs", dict(type=int, default=500)),
    ("--clipping_threshold_d", dict(type=int, default=100)),
    ("--clipping_threshold_g", dict(type=int, default=500)),
    ("--num_gen_steps", dict(type=int, default=1)),
    ("--inp_format", dict(choices=["rel", "abs", "abs_rel"], default="rel")),
    ("--keep_gen_steps", dict(type=int, default=0)),
    ("--top_k_test", dict(type=int, default=20)),
    ("--val_every", dict(type=int, default=1)),
    ("--save_every", dict(type=int, default=5)),
    ("--num_unrolling_steps", dict(type=int, default=0)),
    ("--debug", dict(action="store_true")),
    ("--n_social_modules", dict(type=int, default=1)),
    ("--g_lr", dict(type=float, default=1e-3)),
    ("--d_lr", dict(type=float, default=1e-3)),
    ("--sigma", dict(type=float, default=1.0)),
    ("--gan_type", dict(type=str, choices=["probgan", "mgan", "infogan", "gan"], default="mgan")),
    ("--experiment", dict(type=str, choices=["multi_generator", "discrete"], default="multi_generator")),
    ("--pool_type", dict(type=str, default="sways")),
    ("--global_disc", dict(type=int, default=1)),
    ("--unconditional", dict(action="store_true")),
    ("--augment", dict(type=int, default=1)),
    ("--noise_dim", dict(type=int, default=8)),
    ("--h_dim", dict(type=int, default=32)),
    ("--decoder_h_dim", dict(type=int, default=32)),
    ("--num_samples", dict(type=int, default=20)),
    ("--num_expectation_samples", dict(type=int, default=1)),
    ("--weighting_target", dict(type=str, choices=["l2", "disc_scores", "endpoint", "mgan", "ml", "none"],
                                default="ml")),
    ("--l2_loss_type", dict(type=str, default="min_g_z", opt=["none", "min_z", "min_g_z", "min_g_min_z", "mse"])),
    ("--num_gens", dict(type=int, default=1, opt=[2, 3, 4, 5], tunable=True)),
    ("--l2_decay_rate", dict(type=float, default=1, opt=[1, 0.99, 0.9])),
    ("--checkpoint", dict(type=str)),
    # SGHMC flags (dead in the reference's default path, kept for CLI / meta_tags compatibility)
    ("--sghmc_alpha", dict(default=0.01, type=float, dest="sghmc_alpha", opt=[0.1, 0.01, 0.001])),
    ("--g_noise_loss_lambda", dict(default=3e-2, type=float, dest="g_noise_loss_lambda")),
    ("--d_noise_loss_lambda", dict(default=3e-2, type=float, dest="d_noise_loss_lambda")),
    ("--d_hist_loss_lambda", dict(default=1.0, type=float, dest="d_hist_loss_lambda")),
    # NS original GAN (non-saturating), MM min-max, W Wasserstein, LS least squares
    ("--gan_obj", dict(default="NS", type=str, dest="gan_obj", opt=["NS", "MM", "LS", "W"])),
]

# additive flags of the B200 implementation
_EXTRA_FLAGS = [
    ("--scene_dim", dict(type=int, default=64, choices=[0, 64],
                         help="0 builds G and D without the scene CNN (img=None); the reference hard-codes 64")),
    ("--synthetic_scenes", dict(type=int, default=64, help="scenes per epoch for the synthetic_* datasets")),
    ("--seed", dict(type=int, default=42)),
    ("--resident_images", dict(type=int, default=0, choices=[0, 1],
                               help="synthetic_* datasets: keep one image per scene resident in HBM and cut the agents' "
                                    "33x33 crops on the device (mggan_scene_crop) instead of shipping host-built features")),
    ("--cuda_graph", dict(type=int, default=1, choices=[0, 1],
                          help="replay the training iteration as a CUDA graph for batch structures seen before")),
    ("--graph_cache", dict(type=int, default=64, help="captured iterations kept (LRU), one per batch structure")),
]


def get_parser():
    parser = HyperOptArgumentParser(strategy="grid_search")
    for flag, kw in _FLAGS + _EXTRA_FLAGS:
        kw = dict(kw)
        if "opt" in kw:
            parser.opt_list(flag, options=kw.pop("opt"), tunable=kw.pop("tunable", False), **kw)
        else:
            parser.add_argument(flag, **kw)
    return parser
