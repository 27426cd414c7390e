"""ORACLE helper (this container only): import the UNMODIFIED reference from /root/reference.

The reference's trainer.py / utils.py import skimage, lpips and matplotlib, which are not
installed; those five module names are stubbed in sys.modules.  `.cuda()` is hard-coded at
trainer.py:285,294, so Tensor.cuda is made a no-op for CPU runs.  Nothing here runs on the GPU
box (the reference is not there); tests that need it skip when /root/reference is absent.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("RCOT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "Net_Restormer.py"))


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
        raise RuntimeError("reference not present")
    if REF not in sys.path:
        sys.path.append(REF)
    import importlib

    name = "Net_Restormer"
    # our own drop-in has the same module name; load the reference under an alias
    spec = importlib.util.spec_from_file_location("ref_Net_Restormer", os.path.join(REF, "Net_Restormer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.save_image = lambda *a, **k: None
    return mod


def import_trainer(workdir: str, argv=()):
    """Imports the reference trainer.py verbatim (its main() is guarded) for CPU execution."""
    import torch

    if not available():
        raise RuntimeError("reference not present")
    sk = _stub("skimage")
    skm = _stub("skimage.metrics", peak_signal_noise_ratio=lambda *a, **k: 0.0,
                structural_similarity=lambda *a, **k: 0.0)
    sk.metrics = skm
    _stub("lpips")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    import importlib

    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in ("Net_Restormer", "utils", "trainer", "util") if k in sys.modules}
    try:
        sys.path.insert(0, REF)
        tr = importlib.import_module("trainer")
        net = sys.modules["Net_Restormer"]
    finally:
        sys.path[:] = saved_path
        ref_mods = {k: sys.modules.pop(k) for k in ("Net_Restormer", "utils", "trainer", "util",
                                                     "util.dataset_utils", "util.image_utils",
                                                     "util.degradation_utils") if k in sys.modules}
        sys.modules.update(saved_mods)
    tr.save_image = lambda *a, **k: None
    net.save_image = lambda *a, **k: None
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    tr.opt = tr.parser.parse_args(["--cuda", "", *argv])
    os.makedirs(os.path.join(workdir, "checksample", tr.opt.type), exist_ok=True)
    return tr, net
