#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
for i in 1 2 3 4 5; do timeout 200 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import rcot_b200
rcot_b200.set_hidden_dtype("bf16")
from oracle import train_ref
from rcot_b200.train_step import OTTrainStep
from tests.test_bench_size import _batch, _nets
P, B = 128, 2
Tp, Fp, T_sd, F_sd = _nets(P)
deg, tgt = _batch(11, B, P)
de_id, alpha = torch.tensor([1, 4]), torch.tensor([0.25, 0.7])
step = OTTrainStep(Tp, Fp, "RMSprop")
r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), False, 1e-4)
print("bf16 mode: loss_F %.6f loss_gp %.4f loss_T %.4f loss_mse %.6f" % tuple(r[k].item() for k in ("loss_F", "loss_gp", "loss_T", "loss_mse")))
rcot_b200.set_hidden_dtype("fp32")
Tp, Fp, T_sd, F_sd = _nets(P)
step = OTTrainStep(Tp, Fp, "RMSprop")
r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), False, 1e-4)
print("fp32 mode: loss_F %.6f loss_gp %.4f loss_T %.4f loss_mse %.6f" % tuple(r[k].item() for k in ("loss_F", "loss_gp", "loss_T", "loss_mse")))
PY
done 2>&1 | grep "mode:"
