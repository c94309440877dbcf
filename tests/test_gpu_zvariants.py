"""GPU parity of the flag-reachable variants whose golden vectors were frozen from the reference after this round's GPU
budget was spent (gan_type "gan", pool_type "sgan", experiment "discrete").  Same checks as tests/test_gpu_golden.py::test_training_iterations_objective_variants."""
import pytest

from test_gpu_golden import _run_iterations, injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_training_iterations_late_variants(golden_late_variant, injected, tmp_path, monkeypatch):  # noqa: F811
    import mggan.model.modules.standard_discrete as SD
    monkeypatch.setattr(SD, "get_global_noise", injected.global_noise)     # the discrete generator draws through its own import
    m = golden_late_variant["meta"]
    assert m["gan_type"] == "gan" or m["pool_type"] == "sgan" or m["experiment"] == "discrete"
    _run_iterations(golden_late_variant, injected, tmp_path)
