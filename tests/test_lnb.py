"""LayerNorm-backward epilogue of pm_gemm (the dz GEMM of a block's backward writes dx directly and accumulates dgamma /
dbeta) against the two-kernel form (GEMM -> rcot_ln_bwd), for every C that takes it, TMA and register-staged variants,
with and without the residual term, fp32 and bf16 gather sources."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,K,H,W", [(48, 144, 16, 16), (96, 510, 16, 32), (96, 288, 6, 10), (192, 384, 8, 16), (48, 254, 128, 128)])
@pytest.mark.parametrize("residual", [True, False])
def test_lnb_epilogue_matches_two_kernel_form(cuda_lib, C, K, H, W, residual):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(C + K + H)
    B = 3
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    x, dy, du = r(B, C, H, W) * 1.3 + 0.2, r(B, C, H, W), r(B, K, H, W)
    gamma = (1 + 0.2 * r(C))
    wt = r(K, C, 1, 1) / C ** 0.5                      # the forward conv C -> K; the dgrad pack maps K -> C
    pk = ops.pack_single(wt, "dgrad")
    stats = ops.ln_stats(x)
    srcs = [du]
    if (H * W) % 128 == 0:
        srcs.append(du.bfloat16())
    for src in srcs:
        dz = ops.pm_gemm(src, pk.ptr(0), C)
        dg0, db0 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        want = ops.ln_bwd(dz, x, stats, gamma, dg0, db0, dy=dy if residual else None)
        dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        got = ops.pm_gemm(src, pk.ptr(0), C, residual=dy if residual else None, lnb=(x, stats, gamma, dg1, db1))
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(dg1, dg0, rtol=1e-4, atol=1e-3 * max(1.0, dg0.abs().max().item()))
        torch.testing.assert_close(db1, db0, rtol=1e-4, atol=1e-3 * max(1.0, db0.abs().max().item()))
