"""Host-side autograd routing of the ResNet bridges (no GPU): `functions_conv.Conv1x1Skip` hands the skip branch's gradient to
conv1's input-gradient GEMM (`+aux` epilogue) instead of letting autograd add two materialised gradients
(prototype/model/image_encoder/modified_resnet.py:40-56: `identity = x` next to `self.conv1(x)`).  The C-ABI GEMM is
replaced by a torch stand-in with the same argument contract (operand major-ness, epilogue codes, aux), so what is
checked here is the wiring: which operands reach which GEMM, and that the gradients equal plain fp32 autograd."""
import pytest
import torch

from declip_b200 import functions_conv as C_
from declip_b200 import ops


class _Calls:
    def __init__(self):
        self.log = []


def _fake_gemm(calls):
    def gemm(a, b, *, a_mn_major=False, b_mn_major=False, epilogue=ops.EPI_BF16, alpha=1.0, bias=None, aux=None, **kw):
        A = a.float().t() if a_mn_major else a.float()          # [M, K]
        B = b.float() if b_mn_major else b.float().t()          # [K, N]
        out = alpha * (A @ B)
        if bias is not None:
            out = out + bias
        calls.log.append((int(epilogue), bool(a_mn_major), bool(b_mn_major), aux is not None))
        if epilogue == ops.EPI_BF16_RESID:
            assert aux is not None and aux.dtype == torch.bfloat16 and aux.shape == out.shape
            return (out + aux.float()).bfloat16()
        if epilogue in (ops.EPI_F32, ops.EPI_F32_ATOMIC):
            return out
        assert epilogue == ops.EPI_BF16 and aux is None
        return out.bfloat16()
    return gemm


@pytest.fixture
def fake(monkeypatch):
    calls = _Calls()
    monkeypatch.setattr(ops, "gemm", _fake_gemm(calls))
    monkeypatch.setattr(C_, "weight_shadow", lambda w, pad_rows=0: w.detach().bfloat16())
    return calls


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("use_skip,use_conv", [(True, True), (False, True), (True, False)])
def test_conv1x1_skip_routes_the_identity_gradient_through_the_dgrad_epilogue(fake, use_skip, use_conv):
    torch.manual_seed(0)
    rows, cin, cout = 96, 64, 32
    x = torch.randn(rows, cin).bfloat16().requires_grad_(True)
    w = (torch.randn(cout, cin, 1, 1) * 0.1).requires_grad_(True)
    gy = torch.randn(rows, cout)
    gs = torch.randn(rows, cin)

    y, skip = C_.Conv1x1Skip.apply(x, w)
    assert skip.data_ptr() == x.data_ptr() and y.shape == (rows, cout)       # the skip output is the input itself
    loss = 0
    if use_conv:
        loss = loss + (y.float() * gy).sum()
    if use_skip:
        loss = loss + (skip.float() * gs).sum()
    loss.backward()

    # plain fp32 autograd of the same graph on the bf16-rounded operands
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().bfloat16().float().requires_grad_(True)
    ref = 0
    if use_conv:
        ref = ref + ((xr @ wr.view(cout, cin).t()).bfloat16().float() * gy).sum()      # y leaves the GEMM as bf16
    if use_skip:
        ref = ref + (xr * gs).sum()
    ref.backward()
    assert _cos(x.grad, xr.grad) > 0.9999
    assert (x.grad.float() - xr.grad).abs().max() <= 2e-2 * xr.grad.abs().max()
    if use_conv:
        assert _cos(w.grad, wr.grad.view_as(w)) > 0.9999
        # forward GEMM, ONE input-gradient GEMM (with the skip gradient as its aux operand when there is one), weight gradient
        dgrad = [c for c in fake.log if c[1:3] == (False, True)]
        assert len(dgrad) == 1 and dgrad[0][0] == (ops.EPI_BF16_RESID if use_skip else ops.EPI_BF16) and dgrad[0][3] == use_skip
        assert sum(1 for c in fake.log if c[0] == ops.EPI_F32_ATOMIC and c[1:3] == (True, True)) == 1
    else:
        assert w.grad is None and len(fake.log) == 1                           # only the forward GEMM ran


def test_bottleneck_style_graph_sums_both_branches_once(fake):
    """conv1 -> (nonlinear) -> + identity, and a second consumer of the skip output (the downsample branch of the first
    block of a layer): every gradient that reaches x arrives through Conv1x1Skip's backward, added exactly once."""
    torch.manual_seed(1)
    rows, c = 64, 32
    x = torch.randn(rows, c).bfloat16().requires_grad_(True)
    w1 = (torch.randn(c, c, 1, 1) * 0.2).requires_grad_(True)
    w2 = (torch.randn(c, c, 1, 1) * 0.2).requires_grad_(True)
    y, skip = C_.Conv1x1Skip.apply(x, w1)
    ds = C_.Conv1x1.apply(skip, w2)                       # "downsample" 1x1 conv on the skip branch
    out = torch.relu(y.float()) + ds.float() + 0.5 * skip.float()
    out.square().sum().backward()

    xr = x.detach().float().requires_grad_(True)
    a = w1.detach().bfloat16().float().view(c, c)
    b = w2.detach().bfloat16().float().view(c, c)
    yr = (xr @ a.t()).bfloat16().float()
    # straight-through for the bf16 rounding of y / ds (the product path rounds the same way)
    yr = xr @ a.t() + (yr - xr @ a.t()).detach()
    dsr = xr @ b.t() + ((xr @ b.t()).bfloat16().float() - xr @ b.t()).detach()
    (torch.relu(yr) + dsr + 0.5 * xr).square().sum().backward()
    assert _cos(x.grad, xr.grad) > 0.9995
