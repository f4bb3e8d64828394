"""pytest -m gpu: every C-ABI kernel against its fp32 statement (tests/kernel_checks.py)."""
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _names():
    from tests.kernel_checks import CHECKS
    return list(CHECKS)


@pytest.mark.parametrize('name', _names())
def test_kernel(name):
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')
    from tests.kernel_checks import CHECKS
    res = CHECKS[name]()
    assert res is not None


def test_tensor_core_path_matches_cuda_core_path():
    """The tcgen05 kernels and the plain CUDA-core statement of the same math agree on-device."""
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')
    from contrastive_masked_unet_b200 import ops
    from contrastive_masked_unet_b200._lib import lib
    from tests.kernel_checks import nhwc, _gen, _randn
    g = _gen(123)
    x = nhwc(_randn((2, 128, 20, 12), g))
    wt = _randn((128, 128, 3, 3), g, 0.05)
    dy = nhwc(_randn((2, 128, 20, 12), g))
    wf, wd = ops.pack_conv3x3(wt)
    y_tc, _ = ops.conv3x3_fprop(x, None, wf)
    dx_tc, _ = ops.conv3x3_dgrad(dy, wd, 128)
    dw_tc = ops.conv3x3_wgrad(x, None, dy)
    lib.cmu_debug_set(0, 1)
    try:
        y_cc, _ = ops.conv3x3_fprop(x, None, wf)
        dx_cc, _ = ops.conv3x3_dgrad(dy, wd, 128)
        dw_cc = ops.conv3x3_wgrad(x, None, dy)
    finally:
        lib.cmu_debug_set(0, 0)
    torch.cuda.synchronize()
    assert (y_tc.float() - y_cc.float()).abs().max() <= 2e-2 * y_cc.float().abs().max()
    assert (dx_tc.float() - dx_cc.float()).abs().max() <= 2e-2 * dx_cc.float().abs().max()
    assert (dw_tc - dw_cc).abs().max() <= 1e-3 * dw_cc.abs().max()
