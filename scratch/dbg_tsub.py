import os, sys, torch
sys.path.insert(0, os.getcwd())
from oracle.make_golden import synth_batch
from oracle import restormer_ref as R
from rcot_b200 import ops
from rcot_b200.engine import Tape
import Net_Restormer as N
from rcot_b200.fnet import FnetProgram
from rcot_b200.tnet import TnetProgram
gold = torch.load("tests/golden/rcot_golden.pt", weights_only=False)
torch.manual_seed(0)
T = N.T_net(decoder=True); F = N.F_net(patch_size=32)
T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
Tp = TnetProgram({k: v.detach().cuda() for k, v in T.named_parameters()}, "cuda")
Fp = FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", 32)
deg, tgt = synth_batch(1, 2, 32)
B, P = 2, 32
# oracle
o = gold["T_out"].clone().requires_grad_(True)
loss, rmse = R.transport_loss(o, deg, tgt, R.fnet_forward(F_sd, o), gold["de_id"], 1.0, 10000.0, True)
loss.backward()
dref = o.grad
tape = Tape()
out = Tp.forward(deg.cuda(), tape)
print("out err", (out.cpu() - gold["T_out"]).abs().max().item())
f, dF = Fp.input_grad(out, -1.0 / B)
acc = torch.zeros(4, device="cuda"); gfou = torch.empty_like(out)
ops.cost_stage1(out, deg.cuda(), tgt.cuda(), gold["de_id"].cuda(), gfou, acc)
dout = torch.empty_like(out)
ops.cost_stage2(out, deg.cuda(), tgt.cuda(), gfou, dF, acc, dout, 1.0, 10000.0, float(B*3*P*P))
d = (dout.cpu() - dref)
print("dout max err", d.abs().max().item(), "scale", dref.abs().max().item(), "n big", (d.abs() > 1e-2).sum().item())
idx = (d.abs() > 1e-2).nonzero()
for i in idx[:10]:
    i = tuple(i.tolist()); print(i, dout.cpu()[i].item(), dref[i].item(), (out.cpu()[i]-tgt[i]).item(), (gold["T_out"][i]-tgt[i]).item())
