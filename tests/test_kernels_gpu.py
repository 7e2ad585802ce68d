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
