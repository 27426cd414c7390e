#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
K='regex:gdfn_fwd_kernel|mdta_p1_kernel|pm_gemm_kernel|dw_gate_kernel|dw_plain_kernel|pk_tma_kernel|pk_gemm_kernel'
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2_fused_traffic.csv python scripts/ncu_fused.py > gpurun_out/r2ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gdfn_fwd_kernel|mdta_p1_kernel" -s 4 -c 4 -o gpurun_out/r2_fused_v6 -f python scripts/ncu_fused.py > gpurun_out/r2ncu2.log 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/r2_fused_traffic.csv') if l.startswith('"')))
by={}
for r in rows:
    by.setdefault(r['ID'],{'k':r['Kernel Name'][:60]})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for i,v in by.items():
    print(i, v['k'], 'rd %.1f MB wr %.1f MB %.1f us'%(v.get('dram__bytes_read.sum',0)/1e6 if v.get('dram__bytes_read.sum',0)>1e4 else v.get('dram__bytes_read.sum',0), v.get('dram__bytes_write.sum',0)/1e6 if v.get('dram__bytes_write.sum',0)>1e4 else v.get('dram__bytes_write.sum',0), v.get('gpu__time_duration.sum',0)/1e3))
PY
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2ncu2.log
