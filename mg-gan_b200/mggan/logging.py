"""Minimal stand-in for `test_tube.Experiment` (the reference's logger, used at
mggan/model/train.py:678-690 and mggan/abstract_train.py:27,36,194,201,273).  test_tube is not a
dependency here; this writes the same files in the same places, which is what the checkpoint
loader (`load_from_path`) and scripts/evaluate.py rely on:

    <save_dir>/<name>/version_<v>/meta_tags.csv   (columns key,value)
    <save_dir>/<name>/version_<v>/metrics.csv
"""
import csv
import os


class Experiment:
    def __init__(self, save_dir=None, name="default", debug=False, version=None, **kwargs):
        self.save_dir = str(save_dir) if save_dir is not None else os.getcwd()
        self.name, self.debug = name, debug
        if version is None:
            version = 0
            root = os.path.join(self.save_dir, name)
            if os.path.isdir(root):
                taken = [int(d.split("_")[1]) for d in os.listdir(root) if d.startswith("version_")]
                version = max(taken) + 1 if taken else 0
        self.version = version
        self.tags, self.metrics = {}, []

    def get_data_path(self, exp_name, exp_version):
        path = os.path.join(self.save_dir, exp_name, "version_{}".format(exp_version))
        if not self.debug:
            os.makedirs(path, exist_ok=True)
        return path

    def argparse(self, argparser):
        """Record every parsed flag (test_tube: Experiment.argparse -> meta_tags.csv)."""
        self.tags.update(vars(argparser))
        self.save()

    tag = argparse

    def log(self, metrics_dict, global_step=None):
        row = {k: float(v) for k, v in metrics_dict.items()}
        if global_step is not None:
            row["global_step"] = global_step
        self.metrics.append(row)

    def save(self):
        if self.debug:
            return
        path = self.get_data_path(self.name, self.version)
        if self.tags:
            with open(os.path.join(path, "meta_tags.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["key", "value"])
                for k, v in self.tags.items():
                    w.writerow([k, v])
        if self.metrics:
            keys = sorted({k for r in self.metrics for k in r})
            with open(os.path.join(path, "metrics.csv"), "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=keys)
                w.writeheader()
                w.writerows(self.metrics)
