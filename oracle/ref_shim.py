"""ORACLE helper (test / baseline infrastructure only): import the UNMODIFIED reference.

Where it comes from: $RCOT_REFERENCE, else /root/reference (this container), else oracle/_ref/
(the copy `oracle/build_ref.sh` makes so that the reference travels to the GPU box; git-ignored).

The reference's trainer.py / utils.py import skimage, lpips and matplotlib, which are not
installed; those module names are stubbed in sys.modules (none of them is used by train()).
`.cuda()` is hard-coded at trainer.py:285,294, so for CPU runs Tensor.cuda is made a no-op.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def ref_dir() -> str:
    for cand in (os.environ.get("RCOT_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "Net_Restormer.py")):
            return cand
    return os.path.join(_HERE, "_ref")


REF = ref_dir()


def available() -> bool:
    return os.path.isfile(os.path.join(ref_dir(), "Net_Restormer.py"))


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    return sys.modules[name]


def import_net():
    """Returns the reference's Net_Restormer module with the per-forward PNG dump disabled."""
    if not available():
        raise RuntimeError("reference not present (run oracle/build_ref.sh where /root/reference exists)")
    ref = ref_dir()
    # our own drop-in has the same module name; load the reference under an alias
    spec = importlib.util.spec_from_file_location("ref_Net_Restormer", os.path.join(ref, "Net_Restormer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.save_image = lambda *a, **k: None
    return mod


_CPU_CUDA_PATCHED = False


def import_trainer(workdir: str, argv=(), cuda: bool = False):
    """Imports the reference trainer.py verbatim (its main() is guarded).  cuda=False: CPU execution
    (`--cuda ""`, Tensor.cuda -> identity); cuda=True: the reference as it runs on a GPU."""
    global _CPU_CUDA_PATCHED
    import torch

    if not available():
        raise RuntimeError("reference not present (run oracle/build_ref.sh where /root/reference exists)")
    ref = ref_dir()
    sk = _stub("skimage")
    skm = _stub("skimage.metrics", peak_signal_noise_ratio=lambda *a, **k: 0.0,
                structural_similarity=lambda *a, **k: 0.0)
    sk.metrics = skm
    _stub("lpips")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        import cv2  # noqa: F401
    except Exception:
        _stub("cv2")
    names = ("Net_Restormer", "utils", "trainer", "util", "util.dataset_utils", "util.image_utils",
             "util.degradation_utils")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    try:
        sys.path.insert(0, ref)
        tr = importlib.import_module("trainer")
        net = sys.modules["Net_Restormer"]
    finally:
        sys.path[:] = saved_path
        for k in names:
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
    tr.save_image = lambda *a, **k: None
    net.save_image = lambda *a, **k: None
    if not cuda and not torch.cuda.is_available() and not _CPU_CUDA_PATCHED:
        torch.Tensor.cuda = lambda self, *a, **k: self
        _CPU_CUDA_PATCHED = True
    # (a CPU run on a box that HAS a GPU wraps the call in `cpu_cuda_noop()` instead)
    tr.opt = tr.parser.parse_args(["--cuda", "1" if cuda else "", *argv])
    os.makedirs(os.path.join(workdir, "checksample", tr.opt.type), exist_ok=True)
    return tr, net


class cpu_cuda_noop:
    """Context manager: Tensor.cuda / Module.cuda become identities (a verbatim CPU run of the reference's
    train() on a box that HAS a GPU -- bench.py --impl reference)."""

    def __enter__(self):
        import torch
        self.t = torch.Tensor.cuda
        torch.Tensor.cuda = lambda s, *a, **k: s
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self.t
        return False
