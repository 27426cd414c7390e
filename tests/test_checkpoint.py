"""Checkpoint contract (reference trainer.py:362-371, :100-117; tester.py:54): whole-module pickles
{"epoch", "Tnet", "Fnet"} that interchange with the reference in BOTH directions.

  * ours -> file -> ours: forward equality after the parameters became views of the flat buffer, and
    resume-then-step equals step-without-a-round-trip;
  * reference -> file -> ours: a checkpoint pickled by a pure reference process (oracle/_ref on its path)
    unpickles into the drop-in classes and reproduces the reference's own CPU forward;
  * ours -> file -> reference: the reference process unpickles our checkpoint into ITS classes and its CPU forward
    reproduces our GPU forward.
The reference files travel as oracle/_ref (oracle/build_ref.sh); tests needing them skip when absent.
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

REF_PROC = r"""
import sys, types, os
ref = sys.argv[1]
sys.path.insert(0, ref)
for n in ("skimage", "skimage.metrics", "lpips", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(n, types.ModuleType(n))
import torch
import Net_Restormer as N          # the UNMODIFIED reference module, under its own name
N.save_image = lambda *a, **k: None
mode, path, P = sys.argv[2], sys.argv[3], int(sys.argv[4])
g = torch.Generator().manual_seed(5)
x = torch.rand(1, 3, P, P, generator=g)
if mode == "dump":
    torch.manual_seed(3)
    T = N.T_net(decoder=True); F = N.F_net(patch_size=P)
    with torch.no_grad():
        out, f = T(x), F(x)
    torch.save({"epoch": 7, "Tnet": T, "Fnet": F}, path)
    torch.save({"out": out, "f": f}, path + ".out")
else:
    ck = torch.load(path, weights_only=False)
    T, F = ck["Tnet"].cpu(), ck["Fnet"].cpu()
    assert type(T).__module__ == "Net_Restormer" and hasattr(T, "latent")
    with torch.no_grad():
        out, f = T(x), F(x)
    torch.save({"out": out, "f": f, "epoch": ck["epoch"]}, path + ".out")
"""


def _ref_dir():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference files not present (oracle/build_ref.sh)")
    return ref_shim.ref_dir()


def _x(P):
    return torch.rand(1, 3, P, P, generator=torch.Generator().manual_seed(5))


def test_save_load_forward_and_resume_step(tmp_path):
    import Net_Restormer as N
    import trainer
    from rcot_b200.train_step import OTTrainStep
    P, B = 32, 2
    trainer.opt = trainer.parser.parse_args(["--patch_size", str(P), "--no_dump"])
    torch.manual_seed(0)
    T, F = N.T_net(decoder=True).cuda(), N.F_net(patch_size=P).cuda()
    x = _x(P).cuda()
    with torch.no_grad():
        y0 = T(x).clone()                      # parameters are now views of the flat buffer
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        trainer.save_checkpoint(T, F, 3)
        path = "checkpoint/model_" + str(trainer.opt.type) + "__" + str(trainer.opt.nEpochs) + "_" + str(trainer.opt.sigma) + ".pth"
        ck = torch.load(path, weights_only=False)
    finally:
        os.chdir(cwd)
    assert ck["epoch"] == 3 and "_program" not in ck["Tnet"].__dict__
    T2, F2 = ck["Tnet"].cuda(), ck["Fnet"].cuda()
    assert list(T2.state_dict()) == list(T.state_dict()) and len(T2.state_dict()) == 816
    with torch.no_grad():
        torch.testing.assert_close(T2(x), y0, rtol=1e-5, atol=1e-5)      # same weights; split-K atomics order only
        torch.testing.assert_close(F2(x), F(x), rtol=1e-5, atol=1e-6)
    # resume-then-step == step: one iteration on the original modules and on the reloaded ones (RMSprop state starts
    # at zero in both, as after the reference's --resume which carries no optimizer state)
    g = torch.Generator().manual_seed(1)
    tgt = torch.rand(B, 3, P, P, generator=g).cuda()
    deg = (tgt + 0.1 * torch.randn(B, 3, P, P, generator=g).cuda())
    ids, alpha = torch.tensor([1, 4]).cuda(), torch.tensor([0.3, 0.6]).cuda()
    outs = []
    for (t, f) in ((T, F), (T2, F2)):
        dev = torch.device("cuda", torch.cuda.current_device())
        st = OTTrainStep(t._get_program(dev), f._get_program(dev), "RMSprop")
        r = st.iteration(deg, tgt, ids, alpha, True, 1e-4)
        with torch.no_grad():
            outs.append((torch.stack([r["loss_F"], r["loss_T"], r["loss_mse"]]).cpu(), t(x).clone()))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-7)
    # post-step outputs: RMSprop's first step is sign-like (|dp| = 10*lr whatever |g|), so weights whose gradient is ~0
    # can move the other way under a different atomics order: isolated pixels differ, everything else is tight
    diff = (outs[0][1] - outs[1][1]).abs()
    assert (diff < 5e-3).float().mean() > 0.999 and diff.max() < 3e-2, (diff.max(), (diff >= 5e-3).float().mean())


def test_reference_checkpoint_loads_into_dropin(tmp_path):
    ref = _ref_dir()
    P = 32
    path = str(tmp_path / "ref_ckpt.pth")
    r = subprocess.run([sys.executable, "-c", REF_PROC, ref, "dump", path, str(P)], capture_output=True, text=True,
                       cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import Net_Restormer as N                  # ours: the pickle's class paths resolve to the drop-in classes
    ck = torch.load(path, weights_only=False)
    assert type(ck["Tnet"]) is N.T_net and type(ck["Fnet"]) is N.F_net and ck["epoch"] == 7
    want = torch.load(path + ".out")
    x = _x(P).cuda()
    with torch.no_grad():
        out = ck["Tnet"].cuda()(x).cpu()
        f = ck["Fnet"].cuda()(x).cpu()
    assert torch.allclose(out, want["out"], rtol=1e-3, atol=1e-4), (out - want["out"]).abs().max()
    assert torch.allclose(f.flatten(), want["f"].flatten(), rtol=1e-3, atol=1e-5)
    # the trainer's --resume path: state_dict of the pickled reference modules into fresh drop-in modules
    T2 = N.T_net(decoder=True).cuda()
    T2.load_state_dict(ck["Tnet"].state_dict())
    with torch.no_grad():
        assert torch.allclose(T2(x).cpu(), want["out"], rtol=1e-3, atol=1e-4)


def test_dropin_checkpoint_loads_into_reference(tmp_path):
    ref = _ref_dir()
    import Net_Restormer as N
    P = 32
    torch.manual_seed(11)
    T, F = N.T_net(decoder=True).cuda(), N.F_net(patch_size=P).cuda()
    x = _x(P).cuda()
    with torch.no_grad():
        out, f = T(x).cpu(), F(x).cpu()
    path = str(tmp_path / "ours.pth")
    torch.save({"epoch": 5, "Tnet": T, "Fnet": F}, path)
    r = subprocess.run([sys.executable, "-c", REF_PROC, ref, "load", path, str(P)], capture_output=True, text=True,
                       cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = torch.load(path + ".out")
    assert got["epoch"] == 5
    assert torch.allclose(out, got["out"], rtol=1e-3, atol=1e-4), (out - got["out"]).abs().max()
    assert torch.allclose(f.flatten(), got["f"].flatten(), rtol=1e-3, atol=1e-5)
