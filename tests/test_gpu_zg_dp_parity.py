"""GPU, >= 2 devices: the sharded (NCCL) training iteration equals the single-GPU one (tools/check_dp.py).  Skipped on a
one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_zg_dp_parity.py -m gpu`."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("graph,nccl", [(False, False), (True, False), (False, True), (True, True)])
def test_two_gpu_iterations_equal_one_gpu(graph, nccl):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_dp.py"), "--iters", "4" if graph else "2"]
    if graph:
        cmd.append("--graph")
    if nccl:
        cmd.append("--nccl")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, (out.stdout[-1500:], out.stderr[-1500:])
    rep = json.loads(lines[-1])
    print(rep)
    assert rep["ok"] and rep["max_abs_diff"] <= rep["tol"], rep
    assert rep["exchange"] == ("nccl" if nccl else "peer-memory kernel"), rep
    assert out.returncode == 0, out.stderr[-1500:]
