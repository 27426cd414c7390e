#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
run() { timeout 100 python - <<'PY'
import sys, os, torch
sys.path.insert(0, '.')
import rcot_b200
rcot_b200.set_hidden_dtype("bf16")
from rcot_b200.train_step import OTTrainStep
from tests.test_bench_size import _batch, _nets
P, B = 128, 2
Tp, Fp, T_sd, F_sd = _nets(P)
deg, tgt = _batch(11, B, P)
de_id, alpha = torch.tensor([1, 4]), torch.tensor([0.25, 0.7])
step = OTTrainStep(Tp, Fp, "RMSprop")
r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), False, 1e-4)
print("PK_SPLIT=%s bf16 loss_T %.4f" % (os.environ.get("RCOT_PK_SPLIT", "1"), r["loss_T"].item()))
PY
}
for i in 1 2 3 4 5 6; do RCOT_PK_SPLIT=0 RCOT_LN_BWD_VAR=1 RCOT_DW_BWD_VAR=10 RCOT_DW_GATE1_VAR=10 RCOT_ATTN_FWD_CTAS=148 RCOT_DIRECT_CONV3=0 run 2>&1 | grep loss_T; done
