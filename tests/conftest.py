"""pytest wiring: `gpu` marker, import paths, golden-vector loader."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mg-gan_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfg1_g1_tiny", "cfg2_g4_eth_noimg", "cfg3_g8_sdd_masked"]
# objective variants (gan_obj LS / MM, weighting_target l2 / endpoint / mgan): whole-iteration vectors only
VARIANT_CASES = ["var_ls_l2", "var_mm_endpoint", "var_ns_mgan"]
# variants frozen after the round's GPU budget was spent: checked against the oracle on the CPU, and on the GPU from a
# late-sorting test file (tests/test_gpu_ze_variants.py) so that a surprise there cannot mask the parity tests that ran
LATE_VARIANT_CASES = ["var_gan_plain", "var_sgan_pool", "var_discrete", "var_mse_unroll"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """npz -> nested dict {group: {key: tensor}} with python seq_start_end."""
    import torch

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for key in z.files:
        grp, _, rest = key.partition("/")
        v = z[key]
        scalar = v.ndim == 0 and (grp == "meta" or rest.startswith("metric") or rest.endswith("/step"))
        out.setdefault(grp, {})[rest] = v.item() if scalar else torch.from_numpy(np.asarray(v))
    out["meta"]["seq_start_end"] = [[int(a), int(b)] for a, b in out["meta"]["seq_start_end"].tolist()]
    return out


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=VARIANT_CASES)
def golden_variant(request):
    return load_golden(request.param)


@pytest.fixture(params=LATE_VARIANT_CASES)
def golden_late_variant(request):
    return load_golden(request.param)
