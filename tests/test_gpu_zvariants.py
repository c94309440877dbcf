"""GPU parity of the flag-reachable variants whose golden vectors were frozen from the reference after this round's GPU
budget was spent (gan_type "gan").  Same checks as tests/test_gpu_golden.py::test_training_iterations_objective_variants."""
import pytest

from test_gpu_golden import _run_iterations, injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_training_iterations_gan_type_plain(golden_late_variant, injected, tmp_path):  # noqa: F811
    assert golden_late_variant["meta"]["gan_type"] == "gan"
    _run_iterations(golden_late_variant, injected, tmp_path)
