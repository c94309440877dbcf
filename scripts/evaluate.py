#!/usr/bin/env python
"""Sweep every `version_*` run of a model folder over k = 1 .. num_preds-1 predictions and write ADE / FDE / Mode and
Precision / Recall to one CSV -- the command line, file name and columns of the reference's scripts/evaluate.py:19-169.

Predictions come from the B200 kernels (`PiNetMultiGeneratorGAN.get_predictions`, every strategy of `get_predict_func`);
the metrics run on the host like the reference's, or on the device with `--metrics_device cuda`
(`mggan_min_ade_fde`, `mggan_tube_inside`: same results).

Multi-GPU (BASELINE.json configs[4]): launched under torchrun, e.g.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/evaluate.py <flags>
every rank predicts a contiguous range of the dataset's scenes (`mggan.distributed.shard_items`), the predictions are
gathered to rank 0 in dataset order (`gather_predictions`) and rank 0 alone scores them and writes the CSV -- the same file a
single process writes (scenes are independent units; the Precision / Recall grouping runs on the gathered set).
"""
import argparse
import collections
import os
import pathlib
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mg-gan_b200"))

import pandas as pd  # noqa: E402
import torch  # noqa: E402

from mggan import evaluation  # noqa: E402
from mggan.data_utils.data_loaders import get_dataloader  # noqa: E402
from mggan.model.train import PiNetMultiGeneratorGAN  # noqa: E402

STRATEGIES = ("sampling", "expected", "smart_expected", "rejection")
SWEEP_ORDER = ("smart_expected", "expected", "sampling")          # what `--pred_strat all` runs, in the reference's order

# flag -> argparse keywords (reference scripts/evaluate.py:19-69; the last two are additive)
FLAGS = {
    "--split": dict(default="all", choices=["upper", "lower", "all"]),
    "--device": dict(default="cuda"),
    "--radius": dict(default=3.0, type=float, help="tube radius of the Precision / Recall test"),
    "--model_path": dict(help="folder holding the version_* directories to evaluate"),
    "--output_folder": dict(required=True),
    "--checkpoint": dict(default="best", help="epoch number or 'best'"),
    "--phase": dict(default="test", choices=["train", "val", "test"]),
    "--eval_set": dict(default=None, help="evaluate on this dataset instead of the training one"),
    "--num_preds": dict(default=20, type=int),
    "--pred_strat": dict(default="all", choices=["all", *STRATEGIES]),
    "--no-precision-recall": dict(action="store_true"),
    "--num_scenes": dict(default=64, type=int, help="synthetic datasets: scenes to evaluate"),
    "--metrics_device": dict(default="host", choices=["host", "cuda"],
                             help="host: numpy metrics like the reference; cuda: the device kernels (same results)"),
}
parser = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
for _flag, _kw in FLAGS.items():
    parser.add_argument(_flag, **_kw)


def applicable(config, strategy):
    """Which strategies make sense for a run (reference :119-128)."""
    if config.num_gens == 1:
        return strategy in ("sampling", "rejection")
    if strategy == "rejection":
        return False
    return not (config.weighting_target == "none" and "smart" in strategy)


def describe(config, strategy):
    """The descriptive CSV columns of one (run, strategy) row (reference :137-152)."""
    return {
        "Model": config.name, "# Generators": config.num_gens, "Decoder dim": config.decoder_h_dim,
        "Generator params": getattr(config, "num_gen_parameters", None), "Prediction strategy": strategy,
        "Mode": config.experiment, "Use Classifier": config.gan_type, "Prior": config.weighting_target,
        "Dataset": config.dataset, "Maximization Samples": config.num_samples,
        "Expectation Samples": config.num_expectation_samples, "L2 loss weight": config.l2_loss_weight,
        "Clf loss weight": config.clf_loss_weight, "Sigma": config.sigma,
    }


def load_run(version_dir, checkpoint):
    try:
        return PiNetMultiGeneratorGAN.load_from_path(version_dir, checkpoint)
    except Exception as exc:             # like the reference: fall back to the best checkpoint of the run
        print(exc)
        return PiNetMultiGeneratorGAN.load_from_path(version_dir, "best")


def score(dataset, preds, ks, args):
    on_device = args.metrics_device == "cuda"
    ade_fde = evaluation.evaluate_ade_fde_cuda if on_device else evaluation.evaluate_ade_fde
    metrics = dict(ade_fde(dataset, preds, ks))
    if not args.no_precision_recall:
        prec_rec = evaluation.evaluate_precision_recall_cuda if on_device else evaluation.evaluate_precision_recall
        metrics.update(prec_rec(dataset, preds, args.radius, ks))
    return metrics


def predict_sharded(trainer, dataset, ks, strategy, ctx, batch_size=32):
    """This rank's scenes -> predictions of the WHOLE dataset on rank 0 (None elsewhere)."""
    from torch.utils.data import DataLoader, Subset
    from mggan.data_utils.data_loaders import seq_collate_scene
    from mggan.distributed import gather_predictions, shard_items
    lo, hi = shard_items(len(dataset), ctx.world_size, ctx.rank)
    loader = DataLoader(Subset(dataset, range(lo, hi)), batch_size=batch_size, shuffle=False, collate_fn=seq_collate_scene)
    preds = trainer.get_predictions(loader, max(ks), strategy=strategy) if hi > lo else None
    return gather_predictions(preds, ctx.group, ctx.rank, ctx.world_size)


def main(argv=None):
    args = parser.parse_args(argv)
    ctx = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        from mggan.distributed import DistContext
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl" if args.device == "cuda" else "gloo")
        ctx = DistContext()
    ks = list(range(1, args.num_preds))                      # k = 1 .. num_preds-1, like the reference (:77)
    strategies = SWEEP_ORDER if args.pred_strat == "all" else (args.pred_strat,)
    root = pathlib.Path(args.model_path)
    out_dir = pathlib.Path(args.output_folder)
    out_dir.mkdir(parents=True, exist_ok=True)
    csv = out_dir / f"{root.stem}_{args.phase}_{args.checkpoint}_{args.split}_{args.pred_strat}_radius_{args.radius}.csv"
    print(csv)
    torch.set_grad_enabled(False)
    runs = sorted(d for d in root.iterdir() if "version" in d.stem)
    table = collections.defaultdict(list)
    for strategy in strategies:
        for version_dir in runs:
            trainer, config = load_run(version_dir, args.checkpoint)
            if not applicable(config, strategy):
                continue
            trainer.G.eval()
            config.augment = False
            if args.eval_set is not None:
                table["Training dataset"].append(config.dataset)
                config.dataset = args.eval_set
            loader = get_dataloader(config.dataset, args.phase, batch_size=32, split=args.split,
                                    num_scenes=args.num_scenes, with_img=getattr(config, "scene_dim", 64) > 0)
            row = describe(config, strategy)
            if ctx is None:
                preds = trainer.get_predictions(loader, max(ks), strategy=strategy)
            else:
                preds = predict_sharded(trainer, loader.dataset, ks, strategy, ctx)
                if ctx.rank != 0:
                    continue                                 # rank 0 scores and writes
            row.update(score(loader.dataset, preds, ks, args))
            for column, value in row.items():
                table[column].append(value)
            pd.DataFrame(table).to_csv(csv)                  # rewritten after every run, like the reference
    if ctx is not None:
        import torch.distributed as dist
        dist.barrier()
    return csv


if __name__ == "__main__":
    main()
