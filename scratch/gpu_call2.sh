#!/bin/bash
# GPU call 2: parity of fused-GDFN v2 + multi-M pk_gemm, A/B benches, ncu of attn_bwd / gdfn_mid v2 / pk_mm.
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c2_tests.log
tail -3 $OUT/c2_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c2_bench_all.json 2> $OUT/c2_bench_all.err
RCOT_FUSED_GDFN_MID=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > $OUT/c2_bench_nofuse.json 2> $OUT/c2_bench_nofuse.err
RCOT_PK_MM=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > $OUT/c2_bench_nomm.json 2> $OUT/c2_bench_nomm.err
for f in all nofuse nomm; do python - $OUT/c2_bench_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
timeout 300 python scratch/detail_prof.py 60 > $OUT/c2_detail.txt 2>&1
{ for a in "255 128" "127 128" "255 64" "510 32"; do timeout 120 python scratch/gf_one.py $a; done
  for a in "510 96 1" "288 96 1" "254 48 1" "144 48 1"; do timeout 120 python scratch/pk_one.py $a; done
  for a in "96 1" "96 2" "192 4" "384 8" "48 1"; do timeout 120 python scratch/attn_one.py $a; done; } > $OUT/c2_micro.txt 2>&1
cat $OUT/c2_micro.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 3 -c 1 -o $OUT/attnb python scratch/attn_one.py 96 1 > $OUT/c2_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_mm -s 3 -c 1 -o $OUT/pkmm python scratch/pk_one.py 510 96 1 > $OUT/c2_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gdfn_mid -s 3 -c 1 -o $OUT/gf2 python scratch/gf_one.py 255 128 > $OUT/c2_ncu3.log 2>&1
for r in attnb pkmm gf2; do
  { python scripts/ncu_summarize.py $OUT/$r.ncu-rep --src 0 --top 45; python scripts/ncu_stalls.py $OUT/$r.ncu-rep 0; } > $OUT/c2_sum_$r.txt 2>&1
done
rm -f $OUT/pk_ln.ncu-rep $OUT/pk_plain.ncu-rep $OUT/gf.ncu-rep
ls -la $OUT | head -40
