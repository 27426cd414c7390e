"""Data-parallel host logic on CPU, world_size 2 over gloo (127.0.0.1):
 (1) trainer.train_one shards a global batch (images, de_id, the CPU-RNG alpha draws) by rank;
 (2) the gradient decomposition OTTrainStep relies on: batch-mean terms scaled by
     1/B_global, the Fourier term a batch SUM, the RMSE through an all-reduced sum of squares --
     summed over ranks it must equal the single-process global-batch gradient (oracle arithmetic)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import trainer
        from oracle import restormer_ref as R
        torch.manual_seed(0)
        B, P = 4, 32
        g = torch.Generator().manual_seed(1)
        tgt = torch.rand(B, 3, P, P, generator=g)
        deg = tgt + 0.1 * torch.randn(B, 3, P, P, generator=g)
        de_id = torch.tensor([1, 4, 4, 1])
        # ---- (1) sharding in trainer.train_one
        trainer.opt = trainer.parser.parse_args(["--batchSize", str(B), "--pairnum", "100", "--cuda_graph", "0"])

        class Stub:
            class T:
                class ps:
                    flat = torch.zeros(1)

            def iteration(self, degraded, target, ids, alpha, paired, lr):
                return {"deg": degraded, "ids": ids, "alpha": alpha, "paired": paired}

        torch.manual_seed(5)
        r, d_sh, t_sh = trainer.train_one(Stub(), ([["a"] * B, de_id], deg, tgt), 0, 1e-4)
        torch.manual_seed(5)
        alpha_global = torch.rand(B, 1, 1, 1).view(B)
        sl = slice(rank * B // world, (rank + 1) * B // world)
        assert torch.equal(d_sh, deg[sl]) and torch.equal(t_sh, tgt[sl])
        assert torch.equal(r["ids"], de_id[sl]) and torch.equal(r["alpha"], alpha_global[sl]) and r["paired"]
        # ---- (2) gradient decomposition of the T-sub objective with a small stand-in map
        w = (0.1 * torch.randn(3, 3, 3, 3, generator=torch.Generator().manual_seed(2))).requires_grad_(True)
        torch.manual_seed(0)
        import Net_Restormer as N
        F_sd = {k: v.detach() for k, v in N.F_net(patch_size=P).state_dict().items()}
        sigma, Sigma = 1.0, 100.0

        def tmap(x):
            return x + torch.nn.functional.conv2d(x, w, padding=1)

        out = tmap(deg)
        loss, _ = R.transport_loss(out, deg, tgt, R.fnet_forward(F_sd, out), de_id, sigma, Sigma, True)
        g_full = torch.autograd.grad(loss, w)[0]
        # local shard, scaled the way OTTrainStep scales it
        o = tmap(deg[sl])
        res = deg[sl] - o
        ssq = (res.detach() ** 2).sum()
        dist.all_reduce(ssq)
        n_global = float(B * 3 * P * P)
        rmse = torch.sqrt(ssq / n_global)
        local = (-R.fnet_forward(F_sd, o).sum() / B                       # batch mean -> 1/B_global
                 + sigma * ((res * res.detach()).sum() / (n_global * rmse))  # d rmse = res/(N*rmse) . d res
                 + sigma * R.fourier_cost(res, de_id[sl])                 # batch SUM -> unscaled
                 + Sigma * (o - tgt[sl]).abs().sum() / n_global)          # batch mean of |.|
        g_loc = torch.autograd.grad(local, w)[0]
        dist.all_reduce(g_loc)
        err = (g_loc - g_full).abs().max().item() / g_full.abs().max().item()
        assert err < 1e-4, err
        ret[rank] = err
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_decomposition():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
