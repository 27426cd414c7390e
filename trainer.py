"""Drop-in for the reference's ``trainer.py``: same command line (flags, defaults), same
``train(training_data_loader, T_optimizer, F_optimizer, Tnet, Fnet, epoch)`` entry point, same
checkpoint layout -- with the whole adversarial iteration (reference trainer.py:247-346) executed by
``rcot_b200.train_step.OTTrainStep`` on hand-written sm_100a kernels instead of PyTorch autograd.

Reference lines mirrored: CLI :22-58 · main :67-165 · evaluate :179-227 · LR schedule :228-231 ·
train :234-360 · save_checkpoint :362-371.

Additive flags (all optional, defaults keep the reference behaviour): ``--seed`` (the reference draws a
random seed, :79), ``--synthetic N`` (train on N seeded synthetic pairs: no image folders needed),
``--no_dump`` (skip the PNG side effects), ``--single_dir/--deblur_dir/--lowlight_dir`` (read by the
reference's TrainDataset but missing from its parser).  Launched under ``torchrun`` the batch of
every iteration is sharded across ranks and gradients are all-reduced with NCCL.
"""
from __future__ import annotations

import argparse
import glob
import math
import os
import random

import torch

from Net_Restormer import F_net, T_net
from utils import freeze, unfreeze  # noqa: F401  (re-exported like the reference)

parser = argparse.ArgumentParser(description="RCOT trainer (B200-native hot path)")
parser.add_argument("--batchSize", type=int, default=4, help="training batch size")
parser.add_argument("--nEpochs", type=int, default=200, help="number of epochs to train for")
parser.add_argument("--lr", type=float, default=1e-4, help="Learning Rate. Default=1e-4")
parser.add_argument("--step", type=int, default=20, help="LR is multiplied by 0.1 every n epochs")
parser.add_argument("--cuda", default=True, help="Use cuda? (any non-empty string is truthy, as in the reference)")
parser.add_argument("--resume", default=None, type=str, help="Path to resume model (default: none")
parser.add_argument("--start-epoch", default=1, type=int, help="Manual epoch number (useful on restarts)")
parser.add_argument("--threads", type=int, default=0, help="Number of threads for data loader to use")
parser.add_argument("--pretrained", default="", type=str, help="Path to pretrained model (default: none)")
parser.add_argument("--gpus", default="0", type=str, help="gpu ids (default: 0)")
parser.add_argument("--pairnum", default=0, type=int, help="num of paired samples")
parser.add_argument('--de_type', nargs='+', default=['denoise_15', 'denoise_25', 'denoise_50', 'derain', 'dehaze'],
                    help='which type of degradations is training and testing for.')
parser.add_argument('--denoise_dir', type=str, default='data/Train/Denoise/')
parser.add_argument('--derain_dir', type=str, default='data/Train/Derain/')
parser.add_argument('--dehaze_dir', type=str, default='data/Train/Dehaze/')
parser.add_argument("--degset", default="./data/test/derain/Rain100L/input/", type=str, help="degraded data")
parser.add_argument("--tarset", default="./data/test/derain/Rain100L/target/", type=str, help="target data")
parser.add_argument("--Sigma", default=10000, type=float)
parser.add_argument("--sigma", default=1, type=float)
parser.add_argument("--optimizer", default="RMSprop", type=str, help="optimizer type")
parser.add_argument("--type", default="Deraining", type=str, help="to distinguish the ckpt name ")
parser.add_argument('--patch_size', type=int, default=64, help='patchsize of input.')
parser.add_argument('--num_workers', type=int, default=4, help='number of workers.')
parser.add_argument('--data_file_dir', type=str, default='data_dir/')
# ---- additive
parser.add_argument('--single_dir', type=str, default='data/Train/single/')
parser.add_argument('--deblur_dir', type=str, default='data/Train/Deblur/')
parser.add_argument('--lowlight_dir', type=str, default='data/Train/Lowlight/')
parser.add_argument("--seed", type=int, default=None, help="fixed seed (default: random, like the reference)")
parser.add_argument("--synthetic", type=int, default=0, help="train on this many seeded synthetic pairs")
parser.add_argument("--no_dump", action="store_true", help="skip checksample PNG dumps")
parser.add_argument("--max_iters", type=int, default=0, help="stop each epoch after this many iterations (0 = all)")
parser.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="storage of the blocks' hidden tensors: fp32 (reference-exact class) or bf16 (half the hidden "
                         "traffic and memory; fp32 accumulation, weights and block inputs/outputs; stated tolerance 2e-2)")
parser.add_argument("--device_data", action="store_true",
                    help="assemble the training patches on the GPU (rcot_b200.data: crop / augmentation / uint8-grid "
                         "noise in one kernel per batch) instead of a CPU DataLoader; with --synthetic N")
parser.add_argument("--cuda_graph", type=int, default=-1,
                    help="replay each iteration as one CUDA graph (fixed batch shape: the last partial batch of an epoch "
                         "is dropped): 1 on, 0 off, -1 (default) on when the per-GPU batch is <= 8")

opt = None
DE_IDS = {'denoise_15': 0, 'denoise_25': 1, 'denoise_50': 2, 'derain': 3, 'dehaze': 4, 'deblur': 5, 'lowlight': 6,
          'single': 7}     # util/dataset_utils.py:40


# ---------------------------------------------------------------------------------- data
class SyntheticPairs(torch.utils.data.Dataset):
    """Seeded synthetic (degraded, clean) patches with the item layout of the reference's
    TrainDataset.__getitem__ (util/dataset_utils.py:278): ([name, de_id], degraded, clean).
    denoise_*: uint8-quantised Gaussian noise as util/degradation_utils.py:21-27; derain: sparse bright
    streaks; dehaze: t*clean + A*(1-t); anything else: blurred-ish additive perturbation."""

    def __init__(self, n, patch, de_types, seed=0):
        self.n, self.P, self.seed = n, patch, seed
        self.ids = [DE_IDS[t] for t in de_types]

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        P = self.P
        de_id = self.ids[i % len(self.ids)]
        clean = torch.floor(torch.rand(3, P, P, generator=g) * 255) / 255
        if de_id < 3:
            sig = (15, 25, 50)[de_id]
            deg = torch.floor(torch.clamp(clean * 255 + sig * torch.randn(3, P, P, generator=g), 0, 255)) / 255
        elif de_id == 3:
            streak = (torch.rand(1, P, P, generator=g) > 0.97).float() * (0.4 + 0.4 * torch.rand(1, generator=g))
            deg = torch.clamp(clean + streak, 0, 1)
        elif de_id == 4:
            t = 0.3 + 0.6 * torch.rand(1, generator=g)
            A = 0.7 + 0.3 * torch.rand(1, generator=g)
            deg = clean * t + A * (1 - t)
        else:
            deg = torch.clamp(clean + 0.1 * torch.randn(3, P, P, generator=g), 0, 1)
        return [f"synthetic_{i}", de_id], deg, clean


class DeviceLoader:
    """Loader facade over rcot_b200.data.DeviceTrainData: yields ([names, de_id(global batch)], degraded, clean) with
    the image tensors already on the GPU -- under data parallelism only this rank's shard is assembled (every rank
    draws the same integers from the same seeded RNG)."""

    def __init__(self, data, batch, iters, rank=0, world=1):
        self.data, self.B, self.iters, self.rank, self.world = data, batch, iters, rank, world

    def __len__(self):
        return self.iters

    def __iter__(self):
        B, r, w = self.B, self.rank, self.world
        for _ in range(self.iters):
            draws = self.data.draw(B)
            _, deg, cln = self.data.assemble(draws[r * B // w:(r + 1) * B // w])
            yield [[f"pool_{d[1]}" for d in draws], torch.tensor([d[0] for d in draws])], deg, cln


# ---------------------------------------------------------------------------------- optimizer facade
class EngineOptimizer:
    """What ``main`` hands to ``train`` in place of torch.optim.*: carries the kind and the
    ``param_groups[0]['lr']`` slot the reference's schedule writes to (trainer.py:240-243); the state
    (RMSprop square averages / Adam moments) lives in flat device buffers inside the train step."""

    def __init__(self, kind, lr):
        if kind not in ("RMSprop", "Adam"):
            raise ValueError(f"--optimizer must be RMSprop or Adam, got {kind!r}")
        self.kind = kind
        self.param_groups = [{"lr": lr}]


def _optimizer_kind(optimizer):
    """'RMSprop' / 'Adam' for an EngineOptimizer or for the torch.optim object the reference's main() builds
    (trainer.py:121-126).  A torch.optim instance is accepted as a hyper-parameter carrier: its ``param_groups[0]['lr']``
    is written by the schedule exactly like the reference does, while the update itself runs in the fused flat-buffer
    kernel (the module parameters are views of that buffer, so they change in place).  Non-default hyper-parameters
    would silently differ from what the kernel applies, so they are rejected."""
    kind = getattr(optimizer, "kind", None)
    if kind is not None:
        return kind
    name = type(optimizer).__name__
    d = optimizer.defaults if hasattr(optimizer, "defaults") else {}
    if name == "RMSprop":
        ok = (d.get("alpha", 0.99) == 0.99 and d.get("eps", 1e-8) == 1e-8 and not d.get("momentum", 0)
              and not d.get("centered", False) and not d.get("weight_decay", 0))
    elif name == "Adam":
        ok = (tuple(d.get("betas", (0.9, 0.999))) == (0.9, 0.999) and d.get("eps", 1e-8) == 1e-8
              and not d.get("weight_decay", 0) and not d.get("amsgrad", False))
    else:
        raise TypeError(f"train(): optimizer must be RMSprop or Adam (torch.optim or EngineOptimizer), got {name}")
    if not ok:
        raise ValueError(f"train(): {name} with non-default hyper-parameters is not supported by the fused optimizer kernels")
    return name


_STEPS = {}


def _train_step(Tnet, Fnet, kind):
    from rcot_b200.train_step import OTTrainStep
    key = (id(Tnet), id(Fnet))
    dev = next(Tnet.parameters()).device
    Tp, Fp = Tnet._get_program(dev), Fnet._get_program(dev)
    ent = _STEPS.get(key)
    if ent is None or ent.T is not Tp or ent.F is not Fp:
        ent = OTTrainStep(Tp, Fp, kind, sigma=opt.sigma, Sigma=opt.Sigma)
        _STEPS[key] = ent
    return ent


def _use_graph(per_gpu_batch):
    g = getattr(opt, "cuda_graph", 0)
    g = int(g) if not isinstance(g, bool) else (1 if g else 0)
    return per_gpu_batch <= 8 if g < 0 else bool(g)


def _world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------------- training
def adjust_learning_rate(optimizer, epoch):
    """lr = opt.lr * 0.1 ** (epoch // opt.step)   (reference :228-231)"""
    return opt.lr * (0.1 ** (epoch // opt.step))


def train_one(step, batch, iteration, lr):
    """One iteration on a host batch ([names, de_id], degraded, target); returns device-side losses."""
    ([_, de_id], degraded, target) = batch
    rank, world = _world()
    de_id = torch.as_tensor(de_id)
    B = de_id.shape[0]                              # the GLOBAL batch
    alpha = torch.rand(B, 1, 1, 1).view(B)          # CPU RNG, global batch, like trainer.py:284
    if world > 1:
        if B % world:
            raise ValueError(f"global batch {B} is not divisible by {world} ranks")
        sl = slice(rank * B // world, (rank + 1) * B // world)
        de_id, alpha = de_id[sl], alpha[sl]
        if degraded.shape[0] == B:                  # host loader: every rank holds the global batch
            degraded, target = degraded[sl], target[sl]
        elif degraded.shape[0] != B // world:       # DeviceLoader: already this rank's shard
            raise ValueError(f"batch of {degraded.shape[0]} images for a global batch of {B} on {world} ranks")
    dev = step.T.ps.flat.device
    degraded = degraded.to(dev, non_blocking=True)
    target = target.to(dev, non_blocking=True)
    de_id = de_id.to(dev, non_blocking=True).long()
    alpha = alpha.to(dev, non_blocking=True)
    paired = iteration < opt.pairnum // opt.batchSize
    run = step.iteration_graphed if _use_graph(degraded.shape[0]) else step.iteration
    return run(degraded.contiguous(), target.contiguous(), de_id, alpha, paired, lr), degraded, target


def train(training_data_loader, T_optimizer, F_optimizer, Tnet, Fnet, epoch):
    lr = adjust_learning_rate(F_optimizer, epoch - 1)
    for g in T_optimizer.param_groups:
        g["lr"] = lr / 2
    for g in F_optimizer.param_groups:
        g["lr"] = lr
    rank, _ = _world()
    if rank == 0:
        print("Epoch={}, lr={}".format(epoch, F_optimizer.param_groups[0]["lr"]))
    step = _train_step(Tnet, Fnet, _optimizer_kind(F_optimizer))
    mse, Tloss, Dloss = [], [], []
    for iteration, batch in enumerate(training_data_loader):
        if opt.max_iters and iteration >= opt.max_iters:
            break
        r, degraded, target = train_one(step, batch, iteration, lr)
        if iteration % 10 == 0:
            vals = torch.stack([r["loss_F"], r["loss_T"], r["loss_mse"]]).tolist()
            Dloss.append(vals[0]); Tloss.append(vals[1]); mse.append(vals[2])
            if rank == 0:
                print("Epoch {}({}/{}):Loss_F: {:.5}, Loss_T: {:.5}, Loss_mse: {:.5}".format(
                    epoch, iteration, len(training_data_loader), vals[0], vals[1], vals[2]))
                if not opt.no_dump:
                    from torchvision.utils import save_image
                    d = './checksample/' + opt.type
                    os.makedirs(d, exist_ok=True)
                    save_image(r["out"], d + '/output.png')
                    save_image(degraded, d + '/degraded.png')
                    save_image(target, d + '/target.png')
                    save_image(2 * (degraded - r["out"]), d + '/res.png')
    f = torch.FloatTensor
    return torch.mean(f(mse)), torch.mean(f(Tloss)), torch.mean(f(Dloss))


def PSNR(pred, gt):
    rmse = math.sqrt(((pred - gt) ** 2).mean())
    return 100 if rmse == 0 else 20 * math.log10(1.0 / rmse)


def evaluate(Tnet, deg_list, tar_list):
    """Full-image PSNR over paired folders (reference :179-227); images whose sides are not
    multiples of 8 are skipped (the net needs % 8; the reference tests % 4 and then fails)."""
    if not deg_list:
        return float("nan")
    import numpy as np
    from PIL import Image
    pp, n = 0.0, 0
    with torch.no_grad():
        for deg_name, tar_name in zip(deg_list, tar_list):
            deg = np.array(Image.open(deg_name).convert('RGB'))
            tar = np.array(Image.open(tar_name).convert('RGB'))
            if deg.shape != tar.shape or deg.shape[0] % 8 or deg.shape[1] % 8:
                continue
            x = torch.from_numpy(deg.transpose(2, 0, 1)).float().div(255).unsqueeze(0).cuda()
            y = Tnet(x).squeeze(0).cpu().numpy().transpose(1, 2, 0)
            pp += PSNR(y, tar.astype("float32") / 255)
            n += 1
    return pp / max(n, 1)


def save_checkpoint(Tnet, Fnet, epoch):
    """{"epoch", "Tnet", "Fnet"} whole-module pickles, same path as the reference (:362-371)."""
    path = "checkpoint/" + "model_" + str(opt.type) + "_" + "_" + str(opt.nEpochs) + "_" + str(opt.sigma) + ".pth"
    os.makedirs("checkpoint/", exist_ok=True)
    torch.save({"epoch": epoch, "Tnet": Tnet, "Fnet": Fnet}, path)
    print("Checkpoint saved to {}".format(path))


def main(argv=None):
    global opt
    opt = parser.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank, _ = _world()
    if rank == 0:
        print(opt)
    if not opt.cuda or not torch.cuda.is_available():
        raise Exception("rcot_b200 runs on a B200 only (no CPU path); the reference's CPU mode is the oracle in oracle/")
    import rcot_b200
    rcot_b200.set_hidden_dtype(opt.dtype)
    opt.seed = random.randint(1, 10000) if opt.seed is None else opt.seed
    if world > 1:
        # every rank must build the same weights, shuffle the same way and draw the same alpha stream: the seed
        # rank 0 drew (the reference draws a random one, :79) is broadcast before anything consumes an RNG
        t = torch.tensor([opt.seed], dtype=torch.int64, device="cuda")
        torch.distributed.broadcast(t, 0)
        opt.seed = int(t.item())
    if rank == 0:
        print("Random Seed: ", opt.seed)
    torch.manual_seed(opt.seed)
    random.seed(opt.seed)
    Tnet = T_net(decoder=True).cuda()
    Fnet = F_net(patch_size=opt.patch_size).cuda()
    if opt.resume and os.path.isfile(opt.resume):
        ck = torch.load(opt.resume, weights_only=False)
        opt.start_epoch = ck["epoch"] + 1
        Tnet.load_state_dict(ck["Tnet"].state_dict())
        Fnet.load_state_dict(ck["Fnet"].state_dict())
    if opt.pretrained and os.path.isfile(opt.pretrained):
        w = torch.load(opt.pretrained, weights_only=False)
        Tnet.load_state_dict(w['model'].state_dict())
        Fnet.load_state_dict(w['discr'].state_dict())
    if world > 1:
        # belt and braces after init / --resume / --pretrained: replicas start from rank 0's weights bit for bit
        for net in (Tnet, Fnet):
            prog = net._get_program(torch.device("cuda", torch.cuda.current_device()))
            torch.distributed.broadcast(prog.ps.flat, 0)
            prog.ps.repack()
    T_optimizer = EngineOptimizer(opt.optimizer, opt.lr / 2)
    F_optimizer = EngineOptimizer(opt.optimizer, opt.lr)
    if opt.synthetic:
        train_set = SyntheticPairs(opt.synthetic, opt.patch_size, opt.de_type, opt.seed)
    else:
        try:
            from util.dataset_utils import TrainDataset     # the reference's own data pipeline (out of scope here)
        except ImportError as e:
            raise SystemExit("real-data training needs the reference's util/ package on PYTHONPATH "
                             "(or use --synthetic N): " + str(e))
        train_set = TrainDataset(opt)
    # the reference keeps the last partial batch (:132-135); under data parallelism (batch must divide by the ranks)
    # or CUDA-graph replay (one captured batch shape) it is dropped instead
    loader = torch.utils.data.DataLoader(train_set, num_workers=opt.threads, batch_size=opt.batchSize, shuffle=True,
                                         generator=torch.Generator().manual_seed(opt.seed), pin_memory=True,
                                         drop_last=(world > 1 or _use_graph(opt.batchSize // world)))
    if opt.device_data:
        if not opt.synthetic:
            raise SystemExit("--device_data currently takes its images from --synthetic N (real folders: fill a "
                             "rcot_b200.data.DevicePool with the decoded uint8 images)")
        from rcot_b200.data import DeviceTrainData, synthetic_pool
        pool, samples = synthetic_pool(opt.synthetic, opt.patch_size + 37, opt.patch_size + 52, opt.de_type, opt.seed)
        loader = DeviceLoader(DeviceTrainData(pool, samples, opt.patch_size, opt.seed), opt.batchSize,
                              max(1, opt.synthetic // opt.batchSize), rank, world)
    deg_list, tar_list = sorted(glob.glob(opt.degset + "*")), sorted(glob.glob(opt.tarset + "*"))
    for epoch in range(opt.start_epoch, opt.nEpochs + 1):
        train(loader, T_optimizer, F_optimizer, Tnet, Fnet, epoch)
        if rank == 0:
            p = evaluate(Tnet, deg_list, tar_list)
            os.makedirs("./checksample/" + opt.type, exist_ok=True)
            with open("./checksample/" + opt.type + "/validation_results.txt", "a") as f:
                f.write(f"Patchsize {opt.patch_size} Epoch {epoch}, psnr {p:.4f}, Batchsize {opt.batchSize}\n")
            save_checkpoint(Tnet, Fnet, epoch)
    if world > 1:
        # CUDA graphs that captured NCCL collectives are still alive: destroy_process_group() / interpreter teardown was
        # observed to hang with them (bench.py, 2 GPUs) -- leave without running destructors
        import sys
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
