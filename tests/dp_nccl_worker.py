"""Worker of tests/test_dp_nccl.py (launched with torch.distributed.run, 2 ranks, one B200 each): one OTTrainStep
iteration with the global batch sharded over the ranks vs the same iteration on the whole batch in one process."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import Net_Restormer as N
    from rcot_b200.fnet import FnetProgram
    from rcot_b200.tnet import TnetProgram
    from rcot_b200.train_step import OTTrainStep
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    torch.distributed.init_process_group("nccl")
    P, B = 32, 4

    def nets():
        torch.manual_seed(0)
        T = N.T_net(decoder=True)
        F = N.F_net(patch_size=P)
        return (TnetProgram({k: v.detach().cuda() for k, v in T.named_parameters()}, "cuda"),
                FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", P))

    g = torch.Generator().manual_seed(4)
    tgt = torch.rand(B, 3, P, P, generator=g)
    deg = tgt + 0.1 * torch.randn(B, 3, P, P, generator=g)
    de_id = torch.tensor([1, 4, 3, 0])
    alpha = torch.rand(B, generator=g)
    for paired in (True, False):
        Tp, Fp = nets()
        dp = OTTrainStep(Tp, Fp, "RMSprop")
        assert dp.world == world
        dp.capture = {}
        sl = slice(rank * B // world, (rank + 1) * B // world)
        r = dp.iteration(deg[sl].cuda(), tgt[sl].cuda(), de_id[sl].cuda(), alpha[sl].cuda(), paired, 1e-4)
        Tp1, Fp1 = nets()
        one = OTTrainStep(Tp1, Fp1, "RMSprop", data_parallel=False)
        assert one.world == 1
        one.capture = {}
        r1 = one.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), paired, 1e-4)
        for k, tol in (("loss_F", 1e-5), ("loss_gp", 2e-3), ("loss_T", 2e-3), ("loss_mse", 1e-5)):
            a, b = r[k].item(), r1[k].item()
            assert abs(a - b) <= tol * abs(b) + 1e-7, (k, a, b)
        # F-sub is evaluated at IDENTICAL weights in both runs, but its "fake" input is T(degraded), and T's per-image
        # Grams are split-K sums accumulated with float atomics: T's output is reproducible only to ~1e-6 between ANY two
        # executions (scripts/diag_tfwd.py: 7.5e-7 run to run in one process), whatever the batch split.  A LeakyReLU
        # pre-activation of the potential that is closer to zero than that then lands on the other side in one of the
        # runs -- of the order of one element per pass among the ~1e6 here -- and such a mask flip moves the weight
        # gradients of its layer and of every earlier layer by 1e-5 .. 1e-3 relative (DESIGN.md section 3,
        # scripts/diag_fnet.py).  Measured on 2 B200s over ~30 passes of this worker: whole F gradient 1.8 - 1.9e-5
        # (no flip), 4.1e-5, 8.5e-4 or 1.9e-3 (one flip) -- DISCRETE levels, because the same few borderline elements
        # recur -- under the round-1 and the current kernel settings alike; with its inputs held fixed the critic step
        # repeats to 3e-7 over 60 runs (scripts/diag_fgrad.py), so this is not a race.  The bound therefore allows a
        # flip; what pins the data-parallel ALGEBRA (1/B_global scaling, alpha slices, exact global RMSE, all-reduce) to
        # fp32 summation order is the T gradient below (8e-7 .. 4e-6 in every pass) and tests/test_dp_gloo.py.
        # GP and T-sub are evaluated after the potential's sign-like first RMSprop step(s) (+-10*lr per weight whatever
        # |g|: ~zero gradients flip between any two summation orders), so they carry that step's 1e-3-class noise.
        fc0 = Fp1.ps.offsets["fc.weight"]
        # (GP: 1.5e-4 .. 7e-4 without, 6.7e-3 with a flip in the preceding F-sub pass)
        for k, tol in (("F", 5e-3), ("GP", 3e-2), ("T", 1e-2)):
            a, b = dp.capture[k].double(), one.capture[k].double()
            err = ((a - b).norm() / b.norm()).item()
            if rank == 0:
                print(f"paired={paired} grads {k}: 2 ranks x {B // world} vs 1 x {B}: rel-L2 {err:.2e}")
            assert err < tol, (paired, k, err)
            if k == "F":
                # the fully connected tail sits after every conv mask; its own agreement is a stable 1.4 - 1.6e-4
                # (the 8192-term reductions of linear_wgrad are split differently for 4 and 8 rows)
                err_fc = ((a[fc0:] - b[fc0:]).norm() / b[fc0:].norm()).item()
                if rank == 0:
                    print(f"paired={paired} grads F, fully connected tail ({a.numel() - fc0} values): rel-L2 {err_fc:.2e}")
                assert err_fc < 5e-4, (paired, "F fc tail", err_fc)
        # post-step weights, measured in units of one RMSprop step (10 * lr: the first step of RMSprop is sign-like,
        # +-10*lr per weight whatever |g|).  Two valid fp32 summation orders (2 x 2 images + all-reduce vs 4 images;
        # split-K atomics make even two runs of the SAME build differ) can differ in two ways only:
        #   * a FLIP: a gradient that is zero to within the summation noise changes sign -> the weight differs by 2 steps;
        #   * for F, whose second step (gradient penalty) is no longer sign-like: a sub-step difference proportional to
        #     the 1e-4..1e-3 relative noise of the penalty gradient (it is evaluated after the first, sign-like step).
        # Asserted: no weight differs by more than two flips, flips are rare (< 1 %), and everything else differs by a
        # small fraction of a step on average (< 2 %; measured <= 3.4e-4).  The COUNT of weights that differ by more than
        # 1e-6 is printed only: with one build it was 1.5 %, 2.6 %, 7 %, 25 % and 52 % of the F network on different
        # runs (a mask flip in the penalty step touches every earlier weight by a sub-step amount) -- the round-1 form
        # of this check (count < 5 %) passed or failed by luck.
        for name, a, b, lr in (("T", Tp.ps.flat, Tp1.ps.flat, 5e-5), ("F", Fp.ps.flat, Fp1.ps.flat, 1e-4)):
            step = 10 * lr
            d = (a - b).abs()
            flips = d > 0.5 * step
            flip_frac = flips.float().mean().item()
            rest = d[~flips]
            rest_mean = (rest.mean().item() / step) if rest.numel() else 0.0
            frac = (d > 1e-6).float().mean().item()
            if rank == 0:
                print(f"paired={paired} {name} weights after the step: max diff {d.max().item() / step:.3f} steps, flips {flip_frac:.4%}, "
                      f"mean non-flip diff {rest_mean:.2e} steps, differing by > 1e-6: {frac:.3%}")
            assert d.max().item() <= 2 * 2 * step + 1e-7, (name, d.max().item())
            assert flip_frac < 1e-2, (name, flip_frac)
            assert rest_mean < 2e-2, (name, rest_mean)
        # replicas stay bit-identical
        w = Tp.ps.flat.clone()
        torch.distributed.broadcast(w, 0)
        assert torch.equal(w, Tp.ps.flat)
    torch.distributed.barrier()
    if rank == 0:
        print("DP_NCCL_OK")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
