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
        # F-sub is evaluated at IDENTICAL weights in both runs: sharding + all-reduce must reproduce it to fp32 summation
        # order.  GP and T-sub are evaluated after the potential's sign-like first RMSprop step(s) (+-10*lr per weight
        # whatever |g|: ~zero gradients flip between any two summation orders), so they carry that step's 1e-3-class noise.
        for k, tol in (("F", 5e-5), ("GP", 1e-2), ("T", 1e-2)):
            a, b = dp.capture[k].double(), one.capture[k].double()
            err = ((a - b).norm() / b.norm()).item()
            assert err < tol, (paired, k, err)
            if rank == 0:
                print(f"paired={paired} grads {k}: 2 ranks x {B // world} vs 1 x {B}: rel-L2 {err:.2e}")
        # post-step weights: identical updates up to sign flips of ~zero gradients (RMSprop's first step is sign-like)
        for a, b, lr in ((Tp.ps.flat, Tp1.ps.flat, 5e-5), (Fp.ps.flat, Fp1.ps.flat, 1e-4)):
            d = (a - b).abs()
            # every weight moved by the same steps up to sign flips of ~zero gradients (a flip = 2 * 10*lr per step)
            assert d.max().item() <= 2 * 2 * 10 * lr + 1e-7 and (d > 1e-6).float().mean().item() < 5e-2, \
                (d.max().item(), (d > 1e-6).float().mean().item())
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
