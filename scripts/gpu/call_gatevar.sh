#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for v in 10 11 21 31 41 50 51; do echo "== RCOT_DW_GATE_VAR=$v (gate bwd)"; RCOT_DW_GATE_VAR=$v timeout 120 python scripts/bench_dw.py 2>&1 | grep "gate fwd" | sed 's/gate fwd.*GB.s   //'; done
for v in 10 11 21 30 31 40; do echo "== RCOT_DW_GATE1_VAR=$v (gate fwd)"; RCOT_DW_GATE1_VAR=$v timeout 120 python scripts/bench_dw.py 2>&1 | grep "gate fwd" | sed 's/   gate bwd.*//'; done
} > gpurun_out/r2_gatevar.txt 2>&1
cat gpurun_out/r2_gatevar.txt
