"""GPU parity tests of every C-ABI kernel against torch references (fp64 on TF32-rounded inputs)."""
import pytest
import torch

import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.mark.parametrize("idx", range(len(kc.ALL)))
def test_kernel_check(idx):
    rows = kc.ALL[idx]()
    torch.cuda.synchronize()
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad
    from pe_b200 import native
    native.lib().check_device()


@pytest.mark.parametrize("idx", range(len(kc.PAIRS)))
def test_kernel_check_cta_pairs(idx):
    """The same checks with tcgen05 cta_group::2 forced (64-, 128- and 256-column tiles, stride 2, fused BatchNorm
    statistics / BatchNorm-backward sums / residual epilogues): the production step runs these launches paired."""
    rows = kc.PAIRS[idx]()
    torch.cuda.synchronize()
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad
    from pe_b200 import native
    native.lib().check_device()
