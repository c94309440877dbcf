"""GPU parity of the opt-in GEMM variants behind `mggan_linear_*` -- 2: FP32 128 x 64 tile with register prefetch
(csrc/linear.cu `gemm_kernel_v2`), 3: tcgen05 tensor cores, 3 x TF32 (csrc/linear_tc.cu) -- on the dense-layer cases of
tests/test_gpu_kernels.py, larger shapes of the bench workload, and a whole reference-frozen training iteration.  Both
variants were written after this round's GPU budget was spent and neither is the default."""
import pytest
import torch

import test_gpu_golden as TG
import test_gpu_kernels as TK
from conftest import load_golden
from test_gpu_golden import injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[2, 3])
def gemm_v2(request):
    from mggan import cuda_ext
    prev = cuda_ext.set_gemm_variant(request.param)
    try:
        yield
    finally:
        cuda_ext.set_gemm_variant(prev)


@pytest.mark.parametrize("M,Kd,Od,act", [(1, 16, 16, 0), (37, 24, 64, 2), (300, 192, 96, 2), (129, 96, 1, 3), (1000, 128, 16, 1),
                                          (65, 65, 8, 0), (4100, 24, 64, 2), (4099, 64, 32, 0), (2048, 192, 96, 2),
                                          (513, 36, 32, 0), (127, 3, 8, 1), (3000, 64, 65, 0)])
def test_linear_variant2(gemm_v2, M, Kd, Od, act):
    from mggan import kernels
    TK.test_linear(kernels, M, Kd, Od, act)


def test_variants_agree_on_a_bench_shape(gemm_v2):
    from mggan import cuda_ext, kernels as K
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16384, 192, generator=g).cuda()
    w = (torch.randn(96, 192, generator=g) * 0.1).cuda()
    b = torch.randn(96, generator=g).cuda()
    y2 = K.linear(x, w, b, K.ACT_LRELU, 0.2)
    variant = cuda_ext.set_gemm_variant(1)
    y1 = K.linear(x, w, b, K.ACT_LRELU, 0.2)
    cuda_ext.set_gemm_variant(variant)
    assert float((y1 - y2).abs().max()) <= 1e-5 * float(y1.abs().max())


def test_training_iteration_variant2(gemm_v2, injected, tmp_path):  # noqa: F811
    TG._run_iterations(load_golden("cfg3_g8_sdd_masked"), injected, tmp_path)
