"""Direct FP32 kernels for the convolutions that end in three channels (csrc/conv3.cu) against fp64 F.conv2d:
the output conv (Net_Restormer.py:326, with the `+ inp_img` residual), the data gradient of patch_embed (:117) and of
F_net's first layer (:443, 5x5), at ragged sizes (tiles cut by the image border, widths that are not multiples of 4)
and at the benchmarked 128x128."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=1e-4, atol=1e-5):
    got = got.detach().cpu().double()
    err = (got - ref).abs()
    tol = atol * max(1.0, ref.abs().max().item()) + rtol * ref.abs()
    assert (err > tol).sum().item() == 0, f"max err {err.max().item():.3e} at scale {ref.abs().max().item():.3e}"


@pytest.mark.parametrize("B,Cin,H,W,ks", [(2, 96, 128, 128, 3), (3, 96, 24, 40, 3), (1, 20, 17, 30, 3), (2, 64, 128, 128, 5),
                                          (2, 64, 32, 32, 5), (1, 7, 19, 70, 5)])
@pytest.mark.parametrize("dgrad", [False, True])
def test_conv_to3(cuda_lib, B, Cin, H, W, ks, dgrad):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(B + Cin + H + ks)
    x = torch.randn(B, Cin, H, W, generator=g)
    res = torch.randn(B, 3, H, W, generator=g)
    if dgrad:
        w = torch.randn(Cin, 3, ks, ks, generator=g) / (Cin * ks * ks) ** 0.5       # the conv maps 3 -> Cin channels
        ref = F.conv_transpose2d(x.double(), w.double(), padding=ks // 2)          # = its data gradient
        got = ops.conv_to3(x.cuda(), w.cuda(), dgrad=True)
    else:
        w = torch.randn(3, Cin, ks, ks, generator=g) / (Cin * ks * ks) ** 0.5
        ref = F.conv2d(x.double(), w.double(), padding=ks // 2) + res.double()
        got = ops.conv_to3(x.cuda(), w.cuda(), residual=res.cuda())
    _close(got, ref)


@pytest.mark.parametrize("B,Cout,H,W,ks", [(2, 64, 128, 128, 5), (3, 48, 24, 40, 3), (1, 20, 17, 30, 3), (2, 64, 32, 32, 5),
                                           (1, 7, 19, 70, 5), (2, 48, 128, 128, 3)])
@pytest.mark.parametrize("mode", ["plain", "bias_act", "mask"])
def test_conv_from3(cuda_lib, B, Cout, H, W, ks, mode):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(B + Cout + H + ks)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(Cout, 3, ks, ks, generator=g) / (3 * ks * ks) ** 0.5
    bias = torch.randn(Cout, generator=g)
    my = torch.randn(B, Cout, H, W, generator=g)
    ref = F.conv2d(x.double(), w.double(), padding=ks // 2)
    if mode == "bias_act":
        ref = F.leaky_relu(ref + bias.double().view(1, -1, 1, 1), 0.2)
        got = ops.conv_from3(x.cuda(), w.cuda(), bias=bias.cuda(), act=True, slope=0.2)
    elif mode == "mask":
        ref = ref * torch.where(my.double() > 0, 1.0, 0.2)
        got = ops.conv_from3(x.cuda(), w.cuda(), mask_y=my.cuda(), slope=0.2)
    else:
        got = ops.conv_from3(x.cuda(), w.cuda())
    _close(got, ref)


@pytest.mark.parametrize("B,Cm,H,W,ks", [(2, 96, 128, 128, 3), (3, 48, 24, 40, 3), (1, 20, 17, 30, 3), (2, 64, 128, 128, 5),
                                         (2, 64, 32, 32, 5), (1, 7, 19, 70, 5), (2, 40, 72, 16, 5)])
@pytest.mark.parametrize("from3", [False, True])
def test_conv3_wgrad(cuda_lib, B, Cm, H, W, ks, from3):
    """dW of a conv with three channels on one side vs fp64 autograd; accumulates onto a non-zero buffer."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(B + Cm + H + ks + int(from3))
    cin, cout = (3, Cm) if from3 else (Cm, 3)
    x = torch.randn(B, cin, H, W, generator=g).double()
    dy = torch.randn(B, cout, H, W, generator=g).double()
    w = torch.randn(cout, cin, ks, ks, generator=g).double().requires_grad_(True)
    (F.conv2d(x, w, padding=ks // 2) * dy).sum().backward()
    dw0 = torch.randn(cout, cin, ks, ks, generator=g)
    dw = dw0.clone().cuda()
    many, three = (dy, x) if from3 else (x, dy)
    ops.conv3_wgrad(many.float().cuda(), three.float().cuda(), dw, from3=from3)
    _close(dw, dw0.double() + w.grad, rtol=2e-4, atol=2e-5)
