"""ORACLE / baseline tooling (never the product path): run the reference's OWN `trainer.train()`
(/root/reference/trainer.py:234-360, verbatim through oracle.ref_shim) on synthetic batches and time it.

Used by bench.py (`--impl reference`: CPU arm; `gpu_eager_baseline`: the same verbatim code in PyTorch
eager on cuda:0) and scripts/loss_curve.py.  The only deviations from `python trainer.py` are the ones
SURVEY.md 8(c) lists: missing third-party modules stubbed, `save_image` disabled (the PNG dumps at
trainer.py:355-358 and Net_Restormer.py:433), batches come from a list instead of TrainDataset.
"""
from __future__ import annotations

import contextlib
import io
import os
import tempfile
import time

import torch

from . import ref_shim


def build_reference(P, device, seed=0, argv=()):
    """(trainer module, Tnet, Fnet, T_optimizer, F_optimizer) exactly as the reference's main() builds them
    (trainer.py:92-93,121-126): T_net(decoder=True), F_net(patch_size), RMSprop lr/2 and lr."""
    work = tempfile.mkdtemp(prefix="rcot_ref_")
    cuda = torch.device(device).type == "cuda"
    tr, net = ref_shim.import_trainer(work, argv=argv, cuda=cuda)
    torch.manual_seed(seed)
    Tnet = net.T_net(decoder=True)
    Fnet = net.F_net(patch_size=P)
    if cuda:
        Tnet, Fnet = Tnet.cuda(), Fnet.cuda()
    if tr.opt.optimizer == "Adam":
        T_opt = torch.optim.Adam(Tnet.parameters(), lr=tr.opt.lr / 2)
        F_opt = torch.optim.Adam(Fnet.parameters(), lr=tr.opt.lr)
    else:
        T_opt = torch.optim.RMSprop(Tnet.parameters(), lr=tr.opt.lr / 2)
        F_opt = torch.optim.RMSprop(Fnet.parameters(), lr=tr.opt.lr)
    return tr, Tnet, Fnet, T_opt, F_opt, work


def run_train(tr, loader, T_opt, F_opt, Tnet, Fnet, work, epoch=1, quiet=True):
    """One call of the reference's train() over `loader` (a list of ([names, de_id], degraded, target))."""
    cwd = os.getcwd()
    os.chdir(work)
    try:
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            if next(Tnet.parameters()).is_cuda:
                tr.train(loader, T_opt, F_opt, Tnet, Fnet, epoch)
            else:
                with ref_shim.cpu_cuda_noop():
                    tr.train(loader, T_opt, F_opt, Tnet, Fnet, epoch)
        return sink.getvalue()
    finally:
        os.chdir(cwd)


def time_reference(batches, P, B, steps, warmup, device="cpu", tf32=None, budget_s=None, threads=None):
    """images/s of the verbatim train() on `device`.  batches: list of host batches (cycled).
    Returns (rate, steps_timed, seconds).  tf32: None = leave torch defaults (what the reference runs with:
    cudnn.allow_tf32 True, matmul False); True/False forces both switches."""
    if threads:
        torch.set_num_threads(threads)
    argv = ["--batchSize", str(B), "--patch_size", str(P), "--pairnum", "1000000000"]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if tf32 is not None:
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    try:
        tr, Tnet, Fnet, T_opt, F_opt, work = build_reference(P, device, argv=argv)
        cuda = torch.device(device).type == "cuda"
        if cuda:
            torch.backends.cudnn.benchmark = True          # trainer.py:85
        seq = [batches[i % len(batches)] for i in range(warmup + steps)]
        if warmup:
            run_train(tr, seq[:warmup], T_opt, F_opt, Tnet, Fnet, work)
        done, t_used = 0, 0.0
        t_start = time.perf_counter()
        # one train() call per timed step when a wall-clock budget applies (CPU arm), else one call for all
        chunks = [[b] for b in seq[warmup:]] if budget_s else [seq[warmup:]]
        for ch in chunks:
            if cuda:
                torch.cuda.synchronize()
            s = time.perf_counter()
            run_train(tr, ch, T_opt, F_opt, Tnet, Fnet, work)
            if cuda:
                torch.cuda.synchronize()
            e = time.perf_counter()
            t_used += e - s
            done += len(ch)
            if budget_s and e - t_start > budget_s and done >= 1:
                break
        return B * done / t_used, done, t_used
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
