#!/usr/bin/env python
"""Benchmark of the RCOT training hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (sm_100a kernels)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: the reference algorithm
                                                                        (oracle port) on the host cores

A "step" is one full adversarial iteration (F-sub critic step, gradient penalty step, T-sub transport
step with the Fourier-residual cost, three optimizer steps) on one synthetic batch of 128x128 patches,
per-GPU batch 32 (weak scaling: global batch = 32 * N; gradients all-reduced with NCCL).
`value`  = images/s with the batches already resident in HBM.
`e2e`    = images/s through trainer.train_one() from pinned HOST batches (H2D inside the timed region,
           losses read back to the host every step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "training images/sec at 128x128 bs=32"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--terms", type=int, default=3, help="3: bf16x3 split products (fp32-class), 1: bf16 products")
    ap.add_argument("--graph", type=int, default=-1,
                    help="replay each iteration as one CUDA graph: 1 on, 0 off, -1 (default) = on (measured at batch 32: 241.3 vs "
                         "248.9 ms per step; the captured iteration pins ~98 GB at batch 32 and is released before the "
                         "eager profiling steps)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the strong-scaling and c2/c4/c5 side measurements")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="storage of the blocks' hidden tensors for the HEADLINE run (bf16: see rcot_b200.set_hidden_dtype)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


def synth_host_batches(n, B, P, seed=0, pin=True):
    """n pinned host batches shaped like the reference DataLoader's: ([names, de_id], degraded, target).
    Half of each batch is denoise sigma=25 (de_id 1, |F|^2 branch), half derain-like (de_id 3, |F| branch)."""
    out = []
    for i in range(n):
        g = torch.Generator().manual_seed(seed * 7919 + i)
        tgt = torch.floor(torch.rand(B, 3, P, P, generator=g) * 255) / 255
        deg = torch.floor(torch.clamp(tgt * 255 + 25 * torch.randn(B, 3, P, P, generator=g), 0, 255)) / 255
        streak = (torch.rand(B, 1, P, P, generator=g) > 0.97).float() * 0.6
        rain = torch.clamp(tgt + streak, 0, 1)
        de_id = torch.tensor([1 if j % 2 == 0 else 3 for j in range(B)])
        deg = torch.where((de_id == 1).view(B, 1, 1, 1), deg, rain)
        pin = pin and torch.cuda.is_available()
        out.append(([[f"s{i}_{j}" for j in range(B)], de_id],
                    deg.pin_memory() if pin else deg, tgt.pin_memory() if pin else tgt))
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.1)
        except Exception as e:  # NVML unavailable: report it instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_config(P, B, world, graph=False):
    return {"workload": f"one full adversarial iteration (F-sub + GP + T-sub, 3 optimizer steps), {P}x{P} patches, "
                        f"per-GPU batch {B}, paired, de_id mix [1,3]" + (", CUDA-graph replay" if graph else ""),
            "patch": P, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": "working set per step (tens of GB) >> 126 MB L2; no explicit flush"}


def block_algorithmic_bytes(B, P):
    """per TransformerBlock call fwd 5*B*C*H*W*4 + W, bwd 8*B*C*H*W*4 + 2W; sum over the 102
    calls of one T_net forward: sum C*N = 49.35 M elements per 128^2 image, block weights 0.2567 GB."""
    elems = 49.35e6 * (P / 128.0) ** 2
    wbytes = 0.2567e9
    return 13 * B * elems * 4 + 3 * wbytes


# ---------------------------------------------------------------------------------- reference arms
REF_CPU_BATCH = 2      # bounded sample of the bs=32 workload: one CPU iteration at batch 2 costs ~3 s on 16 threads


def cpu_reference_rate(P, B, steps, warmup, budget_s=200.0):
    """The reference's OWN trainer.train() (verbatim, oracle/_ref or /root/reference through oracle.ref_shim) on the
    host cores, PyTorch CPU fp32, all threads.  Falls back to the validated port (oracle/train_ref.py) only when the
    reference files did not travel.  Returns (images/s, cores, steps timed, kind)."""
    from oracle import ref_shim
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batches = synth_host_batches(max(1, min(4, warmup + steps)), B, P, seed=1, pin=False)
    if ref_shim.available():
        from oracle import ref_run
        torch.manual_seed(0)
        rate, done, _ = ref_run.time_reference(batches, P, B, steps, warmup, device="cpu", budget_s=budget_s)
        return rate, cores, done, "reference"
    import Net_Restormer as N
    from oracle import train_ref
    torch.manual_seed(0)
    T = N.T_net(decoder=True)
    F = N.F_net(patch_size=P)
    T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
    F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
    Ts, Fs = {}, {}
    t_used, done, t0 = 0.0, 0, time.perf_counter()
    for i in range(warmup + steps):
        ([_, de_id], deg, tgt) = batches[i % len(batches)]
        alpha = torch.rand(B)
        s = time.perf_counter()
        with torch.enable_grad():
            train_ref.train_iteration(T_sd, F_sd, Ts, Fs, deg, tgt, de_id, alpha, 1e-4, 1.0, 10000.0, True)
        e = time.perf_counter()
        if i >= warmup:
            t_used += e - s
            done += 1
        if e - t0 > budget_s and done >= 1:
            break
    return B * done / t_used, cores, done, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Bs = REF_CPU_BATCH
    W = max(1, min(args.warmup, 1))
    rate, cores, done, kind = cpu_reference_rate(args.patch, Bs, args.steps, W, budget_s=150.0)
    cfg = make_config(args.patch, Bs, 1)
    cfg["workload"] += (f"  [reference arm: bounded sample -- the same iteration at batch {Bs} per step instead of 32, "
                        "verbatim trainer.train() on the host cores]")
    what = ("the reference's own trainer.train() (verbatim, oracle/_ref)" if kind == "reference"
            else "oracle/train_ref.py port (reference files absent)")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": W, "ms_per_step": 1000.0 * Bs / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{done} timed iteration(s) after {W} warm-up of {what} at batch {Bs}, {args.patch}x"
                                       f"{args.patch}, PyTorch CPU fp32, {cores} threads"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eager_baseline(args):
    """Child process of the b200 arm (`--impl eager`): the verbatim reference trainer.train() in PyTorch eager on
    cuda:0 -- the honest GPU competitor (SURVEY 8d "secondary").  Prints one JSON object."""
    from oracle import ref_run, ref_shim
    out = {"what": "verbatim reference trainer.train() (oracle/_ref), PyTorch eager on cuda:0, cudnn.benchmark=True, "
                   "save_image disabled", "patch": args.patch, "unit": UNIT}
    if not ref_shim.available():
        out["unavailable"] = "reference files absent (oracle/build_ref.sh not run)"
        print(json.dumps(out))
        return
    for label, tf32 in (("tf32_default", None), ("tf32_off", False), ("tf32_on", True)):
        B = args.batch
        while B >= 1:
            try:
                batches = synth_host_batches(2, B, args.patch, seed=1, pin=False)
                torch.manual_seed(0)
                rate, done, secs = ref_run.time_reference(batches, args.patch, B, max(3, args.steps), 2, device="cuda",
                                                          tf32=tf32)
                out[label] = {"value": rate, "batch": B, "steps": done, "ms_per_step": 1000.0 * secs / done,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
                break
            except torch.cuda.OutOfMemoryError:
                out.setdefault("oom_at_batch", []).append(B)
                B //= 2
            finally:
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
    out["note"] = ("tf32_default = torch defaults, what `python trainer.py` runs with (cudnn.allow_tf32=True for the convs, "
                   "matmul fp32); tf32_off = both switches off (true fp32); tf32_on = both on")
    print(json.dumps(out))


# ---------------------------------------------------------------------------------- B200 arm
def _finish(world):
    """End of a rank: with CUDA graphs that captured NCCL collectives alive, destroy_process_group() was observed to
    hang (2-GPU run: the JSON line printed, the process never exited).  Nothing of value remains to be torn down, so
    multi-rank processes flush and leave without running the interpreter's / NCCL's destructors."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_b200(args):
    import Net_Restormer as N
    import trainer
    from rcot_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    eager = None
    if world == 1 and not args.no_eager_baseline:
        # the honest GPU competitor, in a child process so that its memory is gone before our arm starts
        import subprocess
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "eager", "--batch", str(args.batch),
                                "--patch", str(args.patch), "--steps", "3"], capture_output=True, text=True, timeout=600)
            js = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            eager = json.loads(js[-1]) if js else {"unavailable": (r.stderr or "no output")[-300:]}
        except Exception as e:
            eager = {"unavailable": f"{type(e).__name__}: {e}"}
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl")
    ops.TERMS = args.terms
    import rcot_b200
    rcot_b200.set_hidden_dtype(args.dtype)
    B, P, K, W = args.batch, args.patch, args.steps, max(args.warmup, 3)
    args.graph = True if args.graph < 0 else bool(args.graph)
    trainer.opt = trainer.parser.parse_args(["--batchSize", str(B * world), "--patch_size", str(P), "--pairnum", "1000000000",
                                             "--no_dump", "--cuda_graph", "1" if args.graph else "0"])
    torch.manual_seed(0)
    Tnet = N.T_net(decoder=True).cuda()
    Fnet = N.F_net(patch_size=P).cuda()
    step = trainer._train_step(Tnet, Fnet, "RMSprop")
    lr = 1e-4
    host = synth_host_batches(4, B, P, seed=rank)
    dev = [(b[1].cuda(), b[2].cuda(), b[0][1].cuda()) for b in host]
    alphas = [torch.rand(B).cuda() for _ in range(4)]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_step(i):
        d, t, ids = dev[i % 4]
        run = step.iteration_graphed if (args.graph and ops.PROF is None and step.timing is None) else step.iteration
        return run(d, t, ids, alphas[i % 4], True, lr)

    for i in range(W):
        resident_step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        r = resident_step(i)
    e1.record()
    barrier()
    launches = ops.LAUNCHES - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = ms.item()
    value = B * world * K / (ms / 1000.0)

    # ---- end to end: pinned host batches -> trainer.train_one -> losses read back every step
    # (host batches hold the GLOBAL batch, identical on every rank; train_one takes this rank's shard)
    host_g = host if world == 1 else synth_host_batches(4, B * world, P, seed=123)
    h2d = sum(t.numel() * t.element_size() for t in (host[0][1], host[0][2], host[0][0][1])) + 4 * B
    torch.manual_seed(1)
    for i in range(2):
        trainer.train_one(step, host_g[i % 4], i, lr)
    barrier()
    e0.record()
    for i in range(K):
        r, _, _ = trainer.train_one(step, host_g[i % 4], i, lr)
        losses = torch.stack([r["loss_F"], r["loss_T"], r["loss_mse"]]).tolist()      # D2H + sync each step
    e1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if not all(map(lambda v: v == v and abs(v) != float("inf"), losses)):
        raise SystemExit(f"bench: non-finite losses in the timed region: {losses}")
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
    e2e = B * world * K / (ms2.item() / 1000.0)

    # ---- per-kernel CUDA-event timing of the same step (one extra step, outside the timed regions)
    roof, kernels = None, None
    peak, peak_src = peaks()
    phases = None
    if args.graph:
        step.release_graphs()       # the eager profiling steps below need the memory the captured iteration pins
        if not args.no_profile:
            step.iteration(*dev[0][:3], alphas[0], True, lr)     # untimed: lets the caching allocator re-grow its pool
            torch.cuda.synchronize()
    if not args.no_profile:
        # phase breakdown of one more step (events only at 7 section boundaries)
        step.timing = []
        resident_step(0)
        torch.cuda.synchronize()
        phases = {k: round(v, 3) for k, v in step.sections_ms().items()}
        step.timing = None
        ops.PROF = ops.Profiler()
        resident_step(0)
        summ = ops.PROF.summary()
        ops.PROF = None
        tot = sum(v["ms"] for v in summ.values())
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4),
                       "GBps": round(v["bytes"] / 1e9 / (v["ms"] / 1e3), 1) if v["ms"] > 0 else None}
                   for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
        top, tv = max(summ.items(), key=lambda kv: kv[1]["ms"])
        achieved = tv["bytes"] / 1e9 / (tv["ms"] / 1e3)
        # DRAM bytes per launch of that kernel from the committed ncu capture of the same workload (if present)
        traffic = None
        import glob
        tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))   # newest capture last (r1, r1b, r2 ...)
        if tps:
            tj = json.load(open(tps[-1]))
            if tj.get("kernel") == top and tj.get("per_gpu_batch") == B and tj.get("patch") == P:
                traffic = tj["dram_bytes_per_launch"]
        roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                "peak_source": peak_src,
                "launches": tv["launches"], "avg_launch_ms": tv["ms"] / tv["launches"],
                "note": "achieved = algorithmic bytes of all launches of this kernel in one step / their summed "
                        "CUDA-event time, taken on one extra step right after the timed region"}
    # ---- side measurements (not the headline): strong scaling at global batch 32 and BASELINE configs c2 / c4 / c5
    def side_run(Bs, paired, ksteps=5, graph=None, Ps=None):
        """ms per step (max over ranks) of the same iteration at per-GPU batch Bs."""
        Ps = P if Ps is None else Ps
        hb = synth_host_batches(2, Bs, Ps, seed=50 + rank, pin=False)
        dv = [(b[1].cuda(), b[2].cuda(), b[0][1].cuda()) for b in hb]
        al = [torch.rand(Bs).cuda() for _ in range(2)]
        use_graph = True if graph is None else graph
        run = step.iteration_graphed if use_graph else step.iteration
        for i in range(3):
            run(dv[i % 2][0], dv[i % 2][1], dv[i % 2][2], al[i % 2], paired, lr)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(ksteps):
            rr = run(dv[i % 2][0], dv[i % 2][1], dv[i % 2][2], al[i % 2], paired, lr)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        fin = bool(torch.isfinite(torch.stack([rr["loss_F"], rr["loss_T"], rr["loss_mse"]])).all().item())
        if use_graph:
            del rr
            step.release_graphs()       # each captured shape pins its own pool (3 GB per image at 128x128)
        return t.item() / ksteps, use_graph, fin

    extras = {}
    if not args.no_extra_configs and B == 32 and P == 128:
        if world > 1 and 32 % world == 0:
            Bs = 32 // world
            try:
                m, gq, fin = side_run(Bs, True)
            except Exception as e:          # e.g. a collective that cannot be captured: measure without the graph
                print(f"bench: strong-scaling side run with CUDA graph failed ({type(e).__name__}: {e}); eager retry",
                      file=sys.stderr)
                m, gq, fin = side_run(Bs, True, graph=False)
            extras["strong"] = {"what": "same iteration, GLOBAL batch fixed at 32 (BASELINE config c3's shape)",
                                "global_batch": 32, "per_gpu_batch": Bs, "value": 32 / (m / 1e3), "unit": UNIT,
                                "ms_per_step": m, "cuda_graph": gq, "scaling": "strong", "finite": fin}
        else:
            extras["strong"] = {"what": "global batch 32 on one GPU = the headline run", "global_batch": 32,
                                "per_gpu_batch": 32, "value": value, "unit": UNIT, "ms_per_step": ms / K,
                                "cuda_graph": args.graph, "scaling": "strong"}
        if world == 1:
            m, gq, fin = side_run(8, True)
            extras["c2"] = {"what": "BASELINE config c2: single GPU, 128x128, batch 8, fp32, paired", "value": 8 / (m / 1e3),
                            "unit": UNIT, "ms_per_step": m, "cuda_graph": gq, "finite": fin}
            m, gq, fin = side_run(2, False)
            extras["c4"] = {"what": "BASELINE config c4 per-GPU share: unpaired (pairnum=0) full G/D alternation with the "
                                    "Fourier-residual cost, 128x128, 2 images per GPU (16 global on 8 GPUs)",
                            "value": 2 / (m / 1e3), "unit": UNIT, "ms_per_step": m, "cuda_graph": gq, "finite": fin}
            m, gq, fin = side_run(4, True)
            extras["c3_shape_fp32"] = {"what": "BASELINE config c3's per-GPU share (4 images, paired) at fp32 storage",
                                       "value": 4 / (m / 1e3), "unit": UNIT, "ms_per_step": m, "cuda_graph": gq, "finite": fin}
            # recompute mode: only the block inputs are kept (6 GB instead of ~98 GB at batch 32); the forward then runs
            # on the one-kernel GDFN / MDTA-phase-1 kernels, the backward recomputes each block's hidden tensors
            step.save_hidden = False
            try:
                torch.cuda.reset_peak_memory_stats()
                m, gq, fin = side_run(32, True)
                extras["recompute_mode"] = {"what": "the headline step (batch 32, paired) keeping only the block inputs for the "
                                                    "backward: fused forward kernels + per-block recompute",
                                            "value": 32 / (m / 1e3), "unit": UNIT, "ms_per_step": m, "cuda_graph": gq,
                                            "finite": fin, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
            finally:
                step.save_hidden = None
            # c5: T_net forward only, 256x256, 8 images per GPU (64 global on 8 GPUs)
            x5 = torch.rand(8, 3, 256, 256, device="cuda")
            with torch.no_grad():
                for _ in range(2):
                    step.T.forward(x5)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    y5 = step.T.forward(x5)
                b.record()
                torch.cuda.synchronize()
            m5 = a.elapsed_time(b) / 5
            fwd_bytes = 5 * 8 * 49.35e6 * 4 * 4 + 0.2567e9
            extras["c5"] = {"what": "BASELINE config c5 per-GPU share: T_net forward only (inference), 256x256, 8 images",
                            "value": 8 / (m5 / 1e3), "unit": "images/s", "ms_per_step": m5,
                            "roofline_blocks_fwd": {"bytes": fwd_bytes, "achieved": fwd_bytes / 1e9 / (m5 / 1e3),
                                                    "frac": fwd_bytes / 1e9 / (m5 / 1e3) / peaks()[0], "unit": "GB/s"},
                            "finite": bool(torch.isfinite(y5).all().item())}
            del x5, y5
            if args.dtype == "fp32":
                # bf16-storage mode of the hidden tensors (BASELINE config c3 names bf16): the headline step and c3's
                # per-GPU share (4 images) again, with pre / qkv / u / g and their gradients stored as bf16
                rcot_b200.set_hidden_dtype("bf16")
                try:
                    torch.cuda.reset_peak_memory_stats()
                    m, gq, fin = side_run(32, True, graph=False)
                    extras["bf16_storage"] = {"what": "the headline step (batch 32, paired) with bf16-stored hidden tensors "
                                                      "on the C <= 96 levels; stated tolerance 2e-2 (tests/test_bf16_mode.py)",
                                              "value": 32 / (m / 1e3), "unit": UNIT, "ms_per_step": m, "finite": fin,
                                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
                    m, gq, fin = side_run(4, True)
                    extras["c3"] = {"what": "BASELINE config c3 per-GPU share: 128x128, 4 images per GPU (32 global on 8 "
                                            "GPUs), paired, bf16-stored hidden tensors", "value": 4 / (m / 1e3), "unit": UNIT,
                                    "ms_per_step": m, "cuda_graph": gq, "finite": fin}
                finally:
                    rcot_b200.set_hidden_dtype("fp32")

    barrier()
    if rank != 0:
        _finish(world)
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": (("fp32 storage, bf16x3 split products on tcgen05 (fp32-class)" if args.terms == 3
                       else "bf16 products, fp32 storage/accumulate") if args.dtype == "fp32" else
                      "bf16-stored hidden tensors (C <= 96 levels), fp32 block I/O, weights and accumulation; stated tolerance 2e-2"),
            "data": "synthetic",
            "config": make_config(P, B, world, args.graph),
            "clocks": sampler.result(), "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms2.item() / K},
            "losses_last_step": losses}
    line.update(extras)
    if roof:
        line["roofline"] = roof
        line["kernels"] = kernels
        line["phases_ms"] = phases
        # the fused MDTA+GDFN target of SURVEY 8(d): 13*B*C*H*W*4 + 3W bytes per block fwd+bwd, summed over the
        # 102 block calls, against the measured T_net forward+backward time (which also holds the 26 glue convs)
        tb = phases["T_forward"] + phases["T_backward"]
        ab = block_algorithmic_bytes(B, P)
        line["roofline_blocks"] = {"bound": "hbm", "what": "T_net forward + backward (102 MDTA+GDFN blocks + glue), "
                                   "algorithmic bytes of the FUSED block target (SURVEY 8d)", "bytes": ab, "ms": round(tb, 3),
                                   "achieved": ab / 1e9 / (tb / 1e3), "peak": peak, "unit": "GB/s",
                                   "frac": ab / 1e9 / (tb / 1e3) / peak}
    if not args.no_cpu_baseline and world == 1:
        try:
            rate, cores, done, kind = cpu_reference_rate(P, REF_CPU_BATCH, 8, 1, budget_s=25.0)
            what = ("the reference's own trainer.train() (verbatim, oracle/_ref)" if kind == "reference"
                    else "oracle/train_ref.py port")
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{done} timed iteration(s) after 1 warm-up of {what} at batch {REF_CPU_BATCH} "
                                              f"(bounded sample of the bs={B} step), {P}x{P}, PyTorch CPU fp32, {cores} threads"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {type(e).__name__}: {e}"}
    if eager is not None:
        line["gpu_eager_baseline"] = eager
        ev = (eager.get("tf32_default") or {}).get("value")
        if ev:
            line["vs_gpu_eager"] = {"ratio_value": value / ev, "ratio_e2e": e2e / ev,
                                    "note": "this arm's images/s / the verbatim reference in PyTorch eager on the same B200 "
                                            "(torch-default TF32 convs), same patch size" +
                                            ("" if eager["tf32_default"]["batch"] == B else
                                             f"; the eager run fits only batch {eager['tf32_default']['batch']}")}
    print(json.dumps(line), flush=True)
    _finish(world)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager":
        run_eager_baseline(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
