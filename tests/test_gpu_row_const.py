"""`bean_row_const_*` (csrc/bean_row_const.cu) against the torch expression it replaced in device_pack.DeviceScreen: the data-only
part of every row's count log-pmf, lgamma(N + 1) - sum lgamma(x + 1) [+ sum x ln(x / max(N, 1))] -- integer counts below and
above the log-factorial table, zeros, empty rows and non-integer entries (which take lgamma)."""
import pytest
import torch

from crispr_bean_b200.device_pack import row_constants

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("with_xlogx", [True, False])
def test_row_constants_equal_the_float64_torch_expression(cuda_device, dtype, with_xlogx):
    g = torch.Generator().manual_seed(3)
    small = torch.poisson(torch.full((2, 3, 500, 4), 40.0), generator=g)
    big = torch.poisson(torch.full((2, 3, 40, 4), 9000.0), generator=g)           # beyond the 4096-entry table: lgamma
    big[0, 0, :5] = torch.tensor([4095.0, 4096.0, 4097.0, 0.0])                   # the table's edge
    frac = small[:, :, :30].clone() + 0.5                                         # non-integer entries: lgamma as well
    zero = torch.zeros((2, 3, 10, 4))
    x = torch.cat([small, big, frac, zero], dim=2).to(dtype).to(cuda_device)
    rc, tot = row_constants(x, with_xlogx)
    x64 = x.double()
    n64 = x64.sum(-1)
    ref = torch.lgamma(n64 + 1) - torch.lgamma(x64 + 1).sum(-1)
    if with_xlogx:
        ref = ref + torch.xlogy(x64, x64 / n64.clamp(min=1.0).unsqueeze(-1)).sum(-1)
    assert torch.equal(tot, n64)
    err = ((rc - ref).abs() / (ref.abs() + 1.0)).max().item()
    # with the x ln(x / N) term a row of 36,000 reads cancels from terms of size N ln N ~ 4e5 to O(10): both evaluations carry
    # ~1e-16 * 4e5 of rounding there (ln x - ln N here, ln(x / N) in torch)
    assert err <= (1e-10 if with_xlogx else 1e-13), err
    assert rc.shape == x.shape[:-1] and rc.dtype == torch.float64


def test_row_constants_refuse_host_tensors():
    from crispr_bean_b200._lib import BeanError

    with pytest.raises(BeanError):
        row_constants(torch.zeros((3, 4)), True)
