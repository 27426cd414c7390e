"""One adversarial iteration of the reference's training loop (trainer.py:247-346) on the engine:

    F-sub   critic step        L_F  = -mean f(target) + mean f(T(x))            RMSprop(lr)   :262-280
    GP      gradient penalty   L_GP = 10 mean_b (|grad f(x~_b)| - 1)^2           RMSprop(lr)   :283-308
    T-sub   transport step     L_T  = -mean f(T(x)) + sigma (rmse + fourier) [+ Sigma L1]      :311-346
                                                                               RMSprop(lr/2)

Value-neutral deviation from the reference: T(x) is evaluated ONCE per iteration (the reference
evaluates it at :271 without a graph and again at :318 with one; T's weights do not change in
between, so both evaluations are identical).

Data parallel: every rank holds a batch shard and the full weights; the three gradient buffers are
all-reduced (sum) before their optimizer steps, and the global-batch RMSE gets its sum of squares
all-reduced between the two cost stages -- with batch-mean terms scaled by 1/B_global and the
Fourier term left as a batch SUM this reproduces the single-process global-batch gradients
(SURVEY section 8e).
"""
from __future__ import annotations

import torch

from . import ops
from .engine import Tape


class FlatOptimizer:
    """torch.optim.RMSprop / Adam (defaults) over a ParamSet's flat buffers."""

    def __init__(self, ps, kind="RMSprop"):
        if kind not in ("RMSprop", "Adam"):
            raise ValueError(f"unsupported optimizer {kind!r} (reference supports RMSprop and Adam)")
        self.ps, self.kind = ps, kind
        self.sq = torch.zeros_like(ps.flat)
        self.m = torch.zeros_like(ps.flat) if kind == "Adam" else None
        # ONE step counter per network.  Known deviation (Adam only): torch.optim.Adam counts steps per parameter, and
        # fc2.bias is skipped in every gradient-penalty step (its gradient is None there, trainer.py:307), so the
        # reference's bias correction for that single scalar uses t/2.  RMSprop (the default) has no step count.
        self.steps = 0

    def step(self, lr, n=None, hyper=None, lr_mult=1.0):
        """``hyper``: device tensor [lr, 1-b1^t, 1-b2^t] -> the step reads its scalars from device memory
        (what a captured CUDA graph replays); otherwise they are kernel arguments."""
        ps = self.ps
        n = ps.numel if n is None else n
        self.steps += 1
        if hyper is not None:
            if self.kind == "RMSprop":
                ops.rmsprop_h(ps.flat, ps.grad, self.sq, n, hyper, lr_mult)
            else:
                ops.adam_h(ps.flat, ps.grad, self.m, self.sq, n, hyper, lr_mult)
        elif self.kind == "RMSprop":
            ops.rmsprop(ps.flat, ps.grad, self.sq, n, lr)
        else:
            ops.adam(ps.flat, ps.grad, self.m, self.sq, n, lr, self.steps)
        ps.repack()

    def bias_corrections(self, step, b1=0.9, b2=0.999):
        return 1.0 - b1 ** step, 1.0 - b2 ** step


class OTTrainStep:
    def __init__(self, Tprog, Fprog, optimizer="RMSprop", sigma=1.0, Sigma=10000.0, group=None, save_hidden=None,
                 data_parallel=True):
        self.T, self.F = Tprog, Fprog
        self.sigma, self.Sigma = float(sigma), float(Sigma)
        self.T_opt = FlatOptimizer(Tprog.ps, optimizer)
        self.F_opt = FlatOptimizer(Fprog.ps, optimizer)
        self.group = group
        self.save_hidden = save_hidden      # None: decide per batch from free HBM
        self.timing = None                  # set to [] to collect (section, start_event, end_event) per iteration
        self._graphs = {}                   # (B, P, paired) -> captured iteration
        self._hyper = None                  # device [3 optimizer steps x (lr, bc1, bc2)] while capturing/replaying
        self.overlap = True                 # data parallel: bucketed T-gradient all-reduce overlapped with the backward
        self._side = None
        self._reduced = set()
        self.capture = None                 # set to {} to keep clones of the three (all-reduced) gradient buffers (tests)
        self.world = 1
        if data_parallel and (group is not None or
                              (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(group)

    def _mark(self, name):
        """Section boundary for bench.py's phase breakdown (CUDA events on the launching stream)."""
        if self.timing is not None and self._hyper is None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timing.append((name, ev))

    def sections_ms(self):
        """{section: ms} from the marks of the last iteration (call after a synchronize)."""
        out = {}
        for (n0, e0), (_, e1) in zip(self.timing[:-1], self.timing[1:]):
            out[n0] = out.get(n0, 0.0) + e0.elapsed_time(e1)
        return out

    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t, group=self.group)

    def _bucket_ready(self, k):
        """Tape marker callback: gradient bucket k of T_net is final -> all-reduce it on the side stream while the
        backward of the remaining modules keeps the SMs busy."""
        a, b = self.T.ps.bucket_ranges[k]
        if b <= a:
            return
        ev = torch.cuda.Event()
        ev.record()
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            torch.distributed.all_reduce(self.T.ps.grad[a:b], group=self.group)
        self._reduced.add(k)

    def iteration(self, degraded, target, de_id, alpha, paired, lr):
        """degraded/target: [B,3,P,P] local shard; de_id: int64 [B]; alpha: [B] gradient-penalty
        interpolation draws (trainer.py:284, CPU RNG in the reference).  Returns device tensors
        {loss_F, loss_gp, loss_T, loss_mse} (global-batch values), no host sync."""
        T, F = self.T, self.F
        B, _, P, _ = degraded.shape
        Bg = B * self.world
        save = self.save_hidden
        if save is None:
            # hidden tensors of all 102 blocks: 15 * 49.35 M floats per 128x128 image (pre, qkv, u, g, mid-block x)
            need = 15.0 * 49.35e6 * 4 * B * (P / 128.0) ** 2
            if ops.HIDDEN_DTYPE == torch.bfloat16:
                need *= 0.58            # levels with C <= 96 hold 84 % of it at half the size
            save = need < 0.55 * torch.cuda.get_device_properties(degraded.device).total_memory
        tape = Tape(save_hidden=bool(save))
        self._mark("T_forward")
        out = T.forward(degraded, tape)
        self._mark("F_critic_step")
        # ---------------- F-sub
        F.ps.zero_grad()
        loss_F = F.critic_step(target, out, Bg)
        self._allreduce(F.ps.grad)
        if self.capture is not None:
            self.capture["F"] = F.ps.grad.clone()
        hy = self._hyper
        self.F_opt.step(lr, hyper=None if hy is None else hy[0:3])
        # ---------------- gradient penalty (on the updated potential)
        self._mark("F_penalty_step")
        F.ps.zero_grad()
        interp = ops.axpby(target, out, a_vec=alpha)
        loss_gp = F.penalty_step(interp, Bg)
        self._allreduce(F.ps.grad)
        if self.capture is not None:
            self.capture["GP"] = F.ps.grad.clone()
        self.F_opt.step(lr, n=F.n_without_fc2_bias,      # fc2.bias has no gradient here -> skipped
                        hyper=None if hy is None else hy[3:6])
        # ---------------- T-sub
        self._mark("T_cost_and_F_input_grad")
        T.ps.zero_grad()
        f, dF = F.input_grad(out, -1.0 / Bg)
        acc = torch.zeros(4, device=out.device)          # [sum res^2, fourier, sum |out-target|, sum f]
        gfou = torch.empty_like(out)
        tgt = target if paired else None
        ops.cost_stage1(out, degraded, tgt, de_id, gfou, acc)
        ops.signed_sum(f, acc[3:], 0, 1.0)
        self._allreduce(acc)
        n_global = float(Bg * 3 * P * P)
        dout = torch.empty_like(out)
        ops.cost_stage2(out, degraded, tgt, gfou, dF, acc, dout, self.sigma, self.Sigma, n_global)
        self._mark("T_backward")
        overlap = self.world > 1 and self.overlap and len(T.ps.bucket_ranges) > 1
        if overlap:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._reduced = set()
            tape.on_marker = self._bucket_ready
        tape.backward(out, dout)
        self._mark("T_allreduce_and_optimizer")
        if overlap:
            for k, (a, b) in enumerate(T.ps.bucket_ranges):       # what no marker covered (the encoder bucket)
                if k not in self._reduced and b > a:
                    self._allreduce(T.ps.grad[a:b])
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
        else:
            self._allreduce(T.ps.grad[:T.ps.n_used])
        if self.capture is not None:
            self.capture["T"] = T.ps.grad.clone()
            self.capture["dout"] = dout.clone()
        self.T_opt.step(lr / 2, n=T.ps.n_used,           # never-used modules have grad None -> skipped
                        hyper=None if hy is None else hy[6:9], lr_mult=0.5)
        self._mark("end")
        rmse = torch.sqrt(acc[0] / n_global)
        loss_T = -acc[3] / Bg + self.sigma * (rmse + acc[1])
        if paired:
            loss_T = loss_T + self.Sigma * acc[2] / n_global
        if self.world > 1:
            self._allreduce(loss_F)
            self._allreduce(loss_gp)
        return {"loss_F": loss_F[0], "loss_gp": loss_gp[0], "loss_T": loss_T, "loss_mse": rmse, "out": out}


    # ------------------------------------------------------------------ CUDA-graph replay
    def iteration_graphed(self, degraded, target, de_id, alpha, paired, lr):
        """Same contract as ``iteration``; the ~8.7 k kernel launches of one iteration are captured once per
        (batch, patch, paired) into a CUDA graph and replayed, so small per-GPU batches are not bound by the
        Python/driver launch rate.  Inputs are copied into static buffers; the learning rate and Adam's bias
        corrections live in a small device tensor refreshed before every replay."""
        B, _, P, _ = degraded.shape
        key = (B, P, bool(paired), ops.HIDDEN_DTYPE)
        ent = self._graphs.get(key)
        dev = degraded.device
        if ent is None:
            st = {"deg": torch.empty_like(degraded), "tgt": torch.empty_like(target),
                  "ids": torch.empty_like(de_id), "alpha": torch.empty_like(alpha),
                  "hyper": torch.zeros(9, device=dev),
                  # the host may run several replays ahead of the device (train() only syncs every 10 iterations):
                  # each replay gets its own pinned staging slot, reused only after the copy out of it has executed
                  "hyper_ring": [torch.zeros(9).pin_memory() for _ in range(8)],
                  "hyper_ev": [None] * 8, "hyper_i": 0}
            st["deg"].copy_(degraded); st["tgt"].copy_(target); st["ids"].copy_(de_id); st["alpha"].copy_(alpha)
            # one eager iteration first: lazy allocations (scratch, saved-tensor caches, kernel attributes)
            snap = self._snapshot()
            self.iteration(st["deg"], st["tgt"], st["ids"], st["alpha"], paired, lr)
            self._restore(snap)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            g = torch.cuda.CUDAGraph()
            self._hyper = st["hyper"]
            n0 = ops.LAUNCHES
            try:
                with torch.cuda.graph(g):
                    out = self.iteration(st["deg"], st["tgt"], st["ids"], st["alpha"], paired, lr)
            finally:
                self._hyper = None
            self._restore(snap)              # capture does not execute, but the step counters advanced
            st["launches"] = ops.LAUNCHES - n0
            ent = self._graphs[key] = (g, st, out)
        g, st, out = ent
        st["deg"].copy_(degraded, non_blocking=True)
        st["tgt"].copy_(target, non_blocking=True)
        st["ids"].copy_(de_id, non_blocking=True)
        st["alpha"].copy_(alpha, non_blocking=True)
        hi = st["hyper_i"]
        st["hyper_i"] = (hi + 1) % len(st["hyper_ring"])
        if st["hyper_ev"][hi] is not None:
            st["hyper_ev"][hi].synchronize()
        h = st["hyper_ring"][hi]
        fs, ts = self.F_opt.steps, self.T_opt.steps
        for slot, (opt, stepno) in enumerate(((self.F_opt, fs + 1), (self.F_opt, fs + 2), (self.T_opt, ts + 1))):
            bc1, bc2 = opt.bias_corrections(stepno)
            h[3 * slot], h[3 * slot + 1], h[3 * slot + 2] = lr, bc1, bc2
        st["hyper"].copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st["hyper_ev"][hi] = ev
        g.replay()
        self.F_opt.steps += 2
        self.T_opt.steps += 1
        ops.LAUNCHES += st["launches"]
        return out

    def release_graphs(self):
        """Drop every captured iteration together with its private memory pool (at batch 32 a captured iteration pins
        ~98 GB of saved hidden tensors: an eager iteration next to it does not fit 180 GB)."""
        import gc
        self._graphs.clear()
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def _snapshot(self):
        """Weights + optimizer state, so the warm-up iteration before a capture leaves no trace."""
        return (self.T.ps.flat.clone(), self.F.ps.flat.clone(), self.T_opt.sq.clone(), self.F_opt.sq.clone(),
                None if self.T_opt.m is None else self.T_opt.m.clone(),
                None if self.F_opt.m is None else self.F_opt.m.clone(), self.T_opt.steps, self.F_opt.steps)

    def _restore(self, snap):
        Tf, Ff, Tsq, Fsq, Tm, Fm, ts, fs = snap
        self.T.ps.flat.copy_(Tf); self.F.ps.flat.copy_(Ff)
        self.T_opt.sq.copy_(Tsq); self.F_opt.sq.copy_(Fsq)
        if Tm is not None:
            self.T_opt.m.copy_(Tm); self.F_opt.m.copy_(Fm)
        self.T_opt.steps, self.F_opt.steps = ts, fs
        self.T.ps.repack(); self.F.ps.repack()
