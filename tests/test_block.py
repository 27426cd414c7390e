"""TransformerBlock (LN + MDTA + LN + GDFN, Net_Restormer.py:201-214) forward and hand-derived
backward on the GPU against the fp64 oracle (oracle/restormer_ref.py + autograd), at every
(C, heads) configuration T_net uses, on small spatial sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _block_params(C, heads, g, dtype=torch.float32):
    hid = int(C * 2.66)
    r = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    return {
        "b.norm1.body.weight": 1 + 0.2 * r(C), "b.norm1.body.bias": 0.2 * r(C),
        "b.attn.temperature": 1 + 0.3 * r(heads, 1, 1),
        "b.attn.qkv.weight": r(3 * C, C, 1, 1) / C ** 0.5,
        "b.attn.qkv_dwconv.weight": r(3 * C, 1, 3, 3) / 3,
        "b.attn.project_out.weight": r(C, C, 1, 1) / C ** 0.5,
        "b.norm2.body.weight": 1 + 0.2 * r(C), "b.norm2.body.bias": 0.2 * r(C),
        "b.ffn.project_in.weight": r(2 * hid, C, 1, 1) / C ** 0.5,
        "b.ffn.dwconv.weight": r(2 * hid, 1, 3, 3) / 3,
        "b.ffn.project_out.weight": r(C, hid, 1, 1) / hid ** 0.5,
    }


def _check(name, got, ref, rtol=1e-3, atol=1e-4):
    got = got.detach().cpu().double().reshape(ref.shape)
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    # gradients are compared relative to the tensor's scale (sums over thousands of pixels)
    tol = atol * max(1.0, scale) + rtol * ref.abs()
    bad = (err > tol).sum().item()
    print(f"{name:32s} max_err={err.max().item():.3e} scale={scale:.3e} bad={bad}/{err.numel()}")
    assert bad == 0, name


@pytest.mark.parametrize("save", [False, True])
@pytest.mark.parametrize("C,heads,H,W", [(48, 1, 16, 16), (96, 1, 8, 16), (96, 2, 8, 8), (96, 4, 8, 8),
                                         (192, 4, 8, 8), (384, 8, 4, 4), (384, 4, 4, 8)])
def test_block_fwd_bwd(cuda_lib, C, heads, H, W, save):
    from oracle import restormer_ref as R
    from rcot_b200 import engine

    g = torch.Generator().manual_seed(C + heads)
    B = 2
    sd = _block_params(C, heads, g)
    x = torch.randn(B, C, H, W, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    y64 = R.transformer_block(x64, sd64, "b.", heads)
    y64.backward(dy.double())

    ps = engine.ParamSet({k: v for k, v in sd.items()}, "cuda")
    bs = engine.BlockSpec(ps, "b.", C, heads)
    ps.finalize()
    tape = engine.Tape(save_hidden=save)   # save: keep hidden tensors; else recompute them in backward
    xd = x.cuda()
    y = engine.block_fwd(bs, xd, tape)
    _check("y", y, y64.detach())
    leaves = tape.backward(y, dy.cuda().clone())
    _check("dx", tape.grad_of(leaves, xd), x64.grad)
    for k in sd:
        _check(k, ps.g[k], sd64[k].grad)


def test_block_grad_accumulates(cuda_lib):
    """Two invocations of the same block (shared weights) sum their weight gradients."""
    from rcot_b200 import engine

    g = torch.Generator().manual_seed(3)
    C, heads = 48, 1
    sd = _block_params(C, heads, g)
    ps = engine.ParamSet(sd, "cuda")
    bs = engine.BlockSpec(ps, "b.", C, heads)
    ps.finalize()
    x = torch.randn(1, C, 8, 8, generator=g).cuda()
    dy = torch.randn(1, C, 8, 8, generator=g).cuda()
    tape = engine.Tape()
    y = engine.block_fwd(bs, x, tape)
    tape.backward(y, dy.clone())
    once = ps.grad.clone()
    tape = engine.Tape()
    y = engine.block_fwd(bs, x, tape)
    tape.backward(y, dy.clone())
    torch.testing.assert_close(ps.grad, 2 * once, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("save", [False, True])
def test_shared_block_invoked_twice_on_one_tape(cuda_lib, save):
    """y = blk(blk(x)) with ONE set of weights (as T_net's decoder modules are used by both passes)."""
    from oracle import restormer_ref as R
    from rcot_b200 import engine

    g = torch.Generator().manual_seed(12)
    C, heads = 96, 2
    sd = _block_params(C, heads, g)
    x = torch.randn(2, C, 8, 8, generator=g)
    dy = torch.randn(2, C, 8, 8, generator=g)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    R.transformer_block(R.transformer_block(x64, sd64, "b.", heads), sd64, "b.", heads).backward(dy.double())
    ps = engine.ParamSet(sd, "cuda")
    bs = engine.BlockSpec(ps, "b.", C, heads)
    ps.finalize()
    tape = engine.Tape(save_hidden=save)
    xd = x.cuda()
    y = engine.block_fwd(bs, engine.block_fwd(bs, xd, tape), tape)
    leaves = tape.backward(y, dy.cuda().clone())
    _check("dx", tape.grad_of(leaves, xd), x64.grad)
    for k in sd:
        _check(k, ps.g[k], sd64[k].grad)


@pytest.mark.parametrize("C,heads,H,W", [(48, 1, 128, 128), (96, 1, 128, 128), (96, 2, 64, 64)])
def test_block_fwd_bwd_at_bench_size(cuda_lib, C, heads, H, W):
    """The shapes that carry 85 % of the bytes of the benchmarked step (level 1: 128x128, level 2: 64x64), hidden
    tensors kept (the mode bench.py runs at batch 32): forward, dX and every dW against the fp64 oracle."""
    from oracle import restormer_ref as R
    from rcot_b200 import engine

    g = torch.Generator().manual_seed(7 * C + heads)
    B = 2
    sd = _block_params(C, heads, g)
    x = torch.randn(B, C, H, W, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    y64 = R.transformer_block(x64, sd64, "b.", heads)
    y64.backward(dy.double())
    ps = engine.ParamSet({k: v for k, v in sd.items()}, "cuda")
    bs = engine.BlockSpec(ps, "b.", C, heads)
    ps.finalize()
    tape = engine.Tape(save_hidden=True)
    xd = x.cuda()
    y = engine.block_fwd(bs, xd, tape)
    _check("y", y, y64.detach())
    leaves = tape.backward(y, dy.cuda().clone())
    _check("dx", tape.grad_of(leaves, xd), x64.grad)
    for k in sd:
        _check(k, ps.g[k], sd64[k].grad)
