#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c3_tests.log
tail -3 $OUT/c3_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c3_bench_all.json 2> $OUT/c3_bench_all.err
RCOT_FUSED_GDFN_MID=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > $OUT/c3_bench_nofuse.json 2> $OUT/c3_bench_nofuse.err
for f in all nofuse; do python - $OUT/c3_bench_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
timeout 300 python scratch/detail_prof.py 60 > $OUT/c3_detail.txt 2>&1
{ for a in "255 128" "127 128" "255 64" "510 32"; do timeout 120 python scratch/gf_one.py $a; done
  for a in "510 96 1" "288 96 1" "254 48 1" "144 48 1" "96 255 0" "96 96 0"; do timeout 120 python scratch/pk_one.py $a; done
  for a in "96 1" "96 2" "192 4" "384 8" "48 1"; do timeout 120 python scratch/attn_one.py $a; done; } > $OUT/c3_micro.txt 2>&1
cat $OUT/c3_micro.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_mm -s 3 -c 1 -o $OUT/pkmm3 python scratch/pk_one.py 510 96 1 > $OUT/c3_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gdfn_mid -s 3 -c 1 -o $OUT/gf3 python scratch/gf_one.py 255 128 > $OUT/c3_ncu3.log 2>&1
for r in pkmm3 gf3; do
  { python scripts/ncu_summarize.py $OUT/$r.ncu-rep --src 0 --top 30; python scripts/ncu_stalls.py $OUT/$r.ncu-rep 0; } > $OUT/c3_sum_$r.txt 2>&1
done
rm -f $OUT/attnb.ncu-rep $OUT/pkmm.ncu-rep $OUT/gf2.ncu-rep
