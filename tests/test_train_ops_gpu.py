"""Each training operator of csrc/k_train.cu against the same op of plain PyTorch (fp32, same device): forward values and
all gradients.  These are the "plain PyTorch fp32 reference of the same op" numerics tests for the training kernels."""
import pytest
import torch
import torch.nn.functional as F

import yfv2  # noqa: F401

pytestmark = pytest.mark.gpu

# the comparison ops must be true fp32: cuDNN / cuBLAS default to TF32 for convolutions, ~3e-4 relative error
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _ops():
    from model import train_ops
    return train_ops


def _cmp(a, b, tol=2e-5):
    scale = max(1.0, float(b.abs().max()))
    assert float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


@pytest.mark.parametrize("N,K,M,H,W,bias", [(3, 24, 24, 11, 9, False), (2, 288, 72, 8, 10, False), (4, 72, 83, 5, 6, True), (2, 96, 96, 22, 22, False)])
def test_conv1x1(N, K, M, H, W, bias):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, K, H, W, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(M, K, 1, 1, device="cuda", generator=g) / K ** 0.5).requires_grad_(True)
    b = torch.randn(M, device="cuda", generator=g).requires_grad_(True) if bias else None
    dy = torch.randn(N, M, H, W, device="cuda", generator=g)
    y = T.Conv1x1.apply(x, w, b)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone() if bias else None)
    x.grad = None; w.grad = None
    if bias: b.grad = None
    yr = F.conv2d(x, w, b)
    yr.backward(dy)
    _cmp(got[0], yr.detach()); _cmp(got[1], x.grad); _cmp(got[2], w.grad, 1e-4)
    if bias: _cmp(got[3], b.grad, 1e-4)


@pytest.mark.parametrize("C,H,W,ks,stride", [(24, 20, 24, 3, 1), (48, 22, 18, 3, 2), (72, 11, 11, 5, 1), (96, 7, 9, 3, 2)])
def test_dwconv(C, H, W, ks, stride):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(3, C, H, W, device="cuda", generator=g, requires_grad=True)
    w = torch.randn(C, 1, ks, ks, device="cuda", generator=g).requires_grad_(True)
    y = T.DwConv.apply(x, w, stride)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone())
    x.grad = None; w.grad = None
    yr = F.conv2d(x, w, None, stride, ks // 2, 1, C)
    yr.backward(dy)
    _cmp(got[0], yr.detach()); _cmp(got[1], x.grad); _cmp(got[2], w.grad, 1e-4)


def test_stem():
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(3, 3, 64, 96, device="cuda", generator=g)
    w = torch.randn(24, 3, 3, 3, device="cuda", generator=g).requires_grad_(True)
    y = T.StemConv.apply(x, w)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    got = (y.detach(), w.grad.clone())
    w.grad = None
    yr = F.conv2d(x, w, None, 2, 1)
    yr.backward(dy)
    _cmp(got[0], yr.detach()); _cmp(got[1], w.grad, 1e-4)


@pytest.mark.parametrize("relu", [False, True])
def test_bn_train(relu):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.randn(5, 48, 9, 7, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = torch.rand(48, device="cuda", generator=g).add_(0.5).requires_grad_(True)
    beta = torch.randn(48, device="cuda", generator=g).requires_grad_(True)
    rm, rv = torch.zeros(48, device="cuda"), torch.ones(48, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    y = T.BnTrain.apply(x, gamma, beta, rm, rv, relu)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), gamma.grad.clone(), beta.grad.clone())
    x.grad = None; gamma.grad = None; beta.grad = None
    yr = F.batch_norm(x, rm2, rv2, gamma, beta, True, 0.1, 1e-5)
    if relu: yr = F.relu(yr)
    yr.backward(dy)
    _cmp(got[0], yr.detach()); _cmp(got[1], x.grad, 1e-4); _cmp(got[2], gamma.grad, 1e-4); _cmp(got[3], beta.grad, 1e-4)
    _cmp(rm, rm2); _cmp(rv, rv2)


def test_maxpool_and_upsample():
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 24, 16, 20, device="cuda", generator=g).relu_().requires_grad_(True)      # ReLU output: many exact ties at 0
    y = T.MaxPool3x3s2.apply(x)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    got = (y.detach(), x.grad.clone()); x.grad = None
    yr = F.max_pool2d(x, 3, 2, 1); yr.backward(dy)
    assert torch.equal(got[0], yr.detach()); _cmp(got[1], x.grad)
    z = torch.randn(2, 5, 7, 3, device="cuda", generator=g).requires_grad_(True)
    u = T.Upsample2x.apply(z)
    du = torch.randn(u.shape, device="cuda", generator=g)
    u.backward(du)
    got = (u.detach(), z.grad.clone()); z.grad = None
    ur = F.interpolate(z, scale_factor=2); ur.backward(du)
    assert torch.equal(got[0], ur.detach()); _cmp(got[1], z.grad)
