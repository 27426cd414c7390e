import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import Net_Restormer as N
from oracle import train_ref
from oracle.make_golden import synth_batch
from rcot_b200.fnet import FnetProgram
from rcot_b200.tnet import TnetProgram
from rcot_b200.train_step import OTTrainStep
P, B, STEPS = 32, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(0)
T = N.T_net(decoder=True); F = N.F_net(patch_size=P)
T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
Tp = TnetProgram({k: v.detach().cuda() for k, v in T.named_parameters()}, "cuda")
Fp = FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", P)
step = OTTrainStep(Tp, Fp, "RMSprop", sigma=1.0, Sigma=10000.0)
Ts, Fs = {}, {}
de_id = torch.tensor([1, 4])
for i in range(STEPS):
    deg, tgt = synth_batch(100 + i, B, P)
    alpha = torch.rand(B, generator=torch.Generator().manual_seed(i))
    r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), True, 1e-4)
    g = [r["loss_F"].item(), r["loss_gp"].item(), r["loss_T"].item(), r["loss_mse"].item()]
    t0 = time.time()
    o = train_ref.train_iteration(T_sd, F_sd, Ts, Fs, deg, tgt, de_id, alpha, 1e-4, 1.0, 10000.0, True)
    c = [o["loss_F"], o["loss_gp"], o["loss_T"], o["loss_mse"]]
    print(i, "gpu", ["%.6g" % v for v in g], "cpu", ["%.6g" % v for v in c], "rel", ["%.2e" % (abs(a - b) / max(abs(b), 1e-12)) for a, b in zip(g, c)], "%.1fs" % (time.time() - t0), flush=True)
