#!/usr/bin/env python
"""Evaluate every `version_*` run under a model folder: ADE / FDE / Mode and Precision / Recall for
k = 1 .. num_preds-1, written to one CSV (reference: scripts/evaluate.py:19-169, same arguments and
output columns).  Prediction runs on the B200 kernels (`PiNetMultiGeneratorGAN.get_predictions`),
metrics on the host.  Every prediction strategy of the reference (`get_predict_func`) runs on the B200 kernels."""
import os
import sys
from argparse import ArgumentParser
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mg-gan_b200"))

import pandas as pd  # noqa: E402
import torch  # noqa: E402

from mggan.data_utils.data_loaders import get_dataloader  # noqa: E402
from mggan.evaluation import (evaluate_ade_fde, evaluate_ade_fde_cuda, evaluate_precision_recall,  # noqa: E402
                              evaluate_precision_recall_cuda)
from mggan.model.train import PiNetMultiGeneratorGAN  # noqa: E402

parser = ArgumentParser()
parser.add_argument("--split", choices=["upper", "lower", "all"], default="all")
parser.add_argument("--device", default="cuda")
parser.add_argument("--radius", type=float, default=3.0)
parser.add_argument("--model_path")
parser.add_argument("--output_folder", required=True)
parser.add_argument("--checkpoint", default="best")
parser.add_argument("--phase", choices=["train", "val", "test"], default="test")
parser.add_argument("--eval_set", default=None)
parser.add_argument("--num_preds", default=20, type=int)
parser.add_argument("--pred_strat", default="all", choices=["all", "sampling", "expected", "smart_expected", "rejection"])
parser.add_argument("--no-precision-recall", action="store_true")
parser.add_argument("--num_scenes", type=int, default=64, help="synthetic datasets: scenes to evaluate")
parser.add_argument("--metrics_device", choices=["host", "cuda"], default="host",
                    help="host: numpy metrics like the reference; cuda: mggan_min_ade_fde / mggan_tube_inside (same results)")

AVAILABLE = ("sampling", "expected", "smart_expected", "rejection")


def main(argv=None):
    args = parser.parse_args(argv)
    num_preds_list = list(range(1, args.num_preds))           # k = 1 .. num_preds-1, like the reference (:77)
    wanted = ["smart_expected", "expected", "sampling"] if args.pred_strat == "all" else [args.pred_strat]
    pred_strats = [s for s in wanted if s in AVAILABLE]
    for s in wanted:
        if s not in AVAILABLE:
            print(f"prediction strategy '{s}' is not on the B200 path yet; skipped")
    model = Path(args.model_path).stem
    out_dir = Path(args.output_folder)
    out_dir.mkdir(parents=True, exist_ok=True)
    output_csv = out_dir / f"{model}_{args.phase}_{args.checkpoint}_{args.split}_{args.pred_strat}_radius_{args.radius}.csv"
    print(output_csv)
    torch.set_grad_enabled(False)
    model_dirs = sorted(d for d in Path(args.model_path).iterdir() if "version" in d.stem)
    all_results = defaultdict(list)
    for pred_strat in pred_strats:
        for model_dir in model_dirs:
            try:
                m, config = PiNetMultiGeneratorGAN.load_from_path(model_dir, args.checkpoint)
            except Exception as e:
                print(e)
                m, config = PiNetMultiGeneratorGAN.load_from_path(model_dir, "best")
            if config.num_gens == 1 and pred_strat not in ("sampling", "rejection"):      # reference :119-123
                continue
            if config.weighting_target == "none" and "smart" in pred_strat:
                continue
            if pred_strat == "rejection" and config.num_gens != 1:
                continue
            m.G.eval()
            config.augment = False
            if args.eval_set is not None:
                all_results["Training dataset"].append(config.dataset)
                config.dataset = args.eval_set
            loader = get_dataloader(config.dataset, args.phase, batch_size=32, split=args.split,
                                    num_scenes=args.num_scenes, with_img=getattr(config, "scene_dim", 64) > 0)
            for col, val in (("Model", config.name), ("# Generators", config.num_gens),
                             ("Decoder dim", config.decoder_h_dim),
                             ("Generator params", getattr(config, "num_gen_parameters", None)),
                             ("Prediction strategy", pred_strat), ("Mode", config.experiment),
                             ("Use Classifier", config.gan_type), ("Prior", config.weighting_target),
                             ("Dataset", config.dataset), ("Maximization Samples", config.num_samples),
                             ("Expectation Samples", config.num_expectation_samples),
                             ("L2 loss weight", config.l2_loss_weight), ("Clf loss weight", config.clf_loss_weight),
                             ("Sigma", config.sigma)):
                all_results[col].append(val)
            preds = m.get_predictions(loader, max(num_preds_list), strategy=pred_strat)
            ade_fde, prec_rec = ((evaluate_ade_fde_cuda, evaluate_precision_recall_cuda) if args.metrics_device == "cuda"
                                 else (evaluate_ade_fde, evaluate_precision_recall))
            metric_dict = dict(ade_fde(loader.dataset, preds, num_preds_list))
            if not args.no_precision_recall:
                metric_dict.update(prec_rec(loader.dataset, preds, args.radius, num_preds_list))
            for k, v in metric_dict.items():
                all_results[k].append(v)
            pd.DataFrame(all_results).to_csv(output_csv)
    return output_csv


if __name__ == "__main__":
    main()
