"""GPU parity tests proper: every kernel family and both model classes, through the C ABI, against the CPU oracle /
golden fixtures.  One pytest case per check group in tests/gpu_checks.py; the failing sub-checks are listed."""
import pytest
import torch

from tests import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device(tdr_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert tdr_lib.tdr_check_device() == 0, tdr_lib.tdr_last_error().decode()


@pytest.mark.parametrize("group", list(gpu_checks.CHECKS))
def test_group(group):
    res = gpu_checks.CHECKS[group]()
    torch.cuda.synchronize()
    bad = [f"{r['name']}: err {r['max_err']} > tol {r.get('tol')} {r.get('note', '')}" for r in res if not r["ok"]]
    assert not bad, "\n".join(bad)


def test_no_cpu_fallback():
    """The product path must refuse CPU tensors instead of silently computing elsewhere."""
    from textualdegremoval_b200 import TdrError, define_network
    net = define_network(dict(type="Restormer", dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1))
    with pytest.raises(TdrError):
        net(torch.rand(1, 3, 64, 64))
