"""GPU: both command lines of the drop-in surface run end to end on a synthetic dataset -- `mggan/model/train.py <flags>`
(reference mggan/model/train.py:665-691: log-dir layout, meta_tags.csv, checkpoints) and `scripts/evaluate.py`
(reference scripts/evaluate.py:72-169: CSV name and columns), the second also sharded over 2 GPUs when the box has them."""
import csv
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, timeout=600):
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, (cmd, out.stdout[-1500:], out.stderr[-3000:])
    return out


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    log = tmp_path_factory.mktemp("cli_logs")
    _run([sys.executable, os.path.join(ROOT, "mg-gan_b200", "mggan", "model", "train.py"), "--dataset", "synthetic_gofp",
          "--synthetic_scenes", "6", "--epochs", "2", "--batch_size", "3", "--num_gens", "2", "--num_samples", "4",
          "--save_every", "1", "--log_dir", str(log), "--name", "cli", "--num_unrolling_steps", "1",
          "--l2_loss_type", "mse"])
    model_dir = log / "multi_generator" / "cli"
    versions = [d for d in model_dir.iterdir() if d.name.startswith("version_")]
    assert len(versions) == 1
    v = versions[0]
    assert (v / "meta_tags.csv").is_file()
    names = sorted(p.name for p in (v / "checkpoints").iterdir())
    # best-checkpoint tracking follows "val/ADE k=20" (abstract_train.py:104-107,184-191): needs the default --top_k_test 20
    assert "checkpoint_best.pth" in names and "checkpoint_1.pth" in names and "checkpoint_2.pth" in names
    ck = torch.load(v / "checkpoints" / "checkpoint_2.pth", map_location="cpu")
    assert set(ck) == {"generator", "discriminator", "gen_opt", "disc_opt"}          # abstract_train.py:236-244
    return model_dir


def _check_csv(path, rows, k_max=3):
    with open(path) as f:
        table = list(csv.DictReader(f))
    assert len(table) == rows
    for row in table:
        for k in range(1, k_max + 1):
            for col in (f"ADE k={k}", f"FDE k={k}", f"Recall k={k}"):
                assert col in row and float(row[col]) == float(row[col]), col
        assert "Precision" in row and row["Prediction strategy"] in ("sampling", "expected", "smart_expected")
        assert float(row["ADE k=3"]) <= float(row["ADE k=1"]) + 1e-9            # min over more samples never gets worse
    return table


def test_train_cli_then_evaluate_cli(trained, tmp_path):
    out = tmp_path / "eval"
    res = _run([sys.executable, os.path.join(ROOT, "scripts", "evaluate.py"), "--model_path", str(trained), "--output_folder",
                str(out), "--checkpoint", "best", "--phase", "test", "--num_preds", "4", "--num_scenes", "6"])
    files = list(out.iterdir())
    assert len(files) == 1 and files[0].name == "cli_test_best_all_all_radius_3.0.csv", files      # reference naming (:83-87)
    assert files[0].name in res.stdout
    _check_csv(files[0], rows=3)             # --pred_strat all: smart_expected, expected, sampling


def test_evaluate_cli_sharded_over_two_gpus(trained, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "eval2"
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
          "--master-port", str(port), os.path.join(ROOT, "scripts", "evaluate.py"), "--model_path", str(trained),
          "--output_folder", str(out), "--checkpoint", "2", "--phase", "test", "--num_preds", "4", "--num_scenes", "6",
          "--pred_strat", "sampling", "--metrics_device", "cuda"])
    files = list(out.iterdir())
    assert len(files) == 1
    _check_csv(files[0], rows=1)
