#!/bin/bash
# GPU call 1 of this session: parity of the new kernels, bench A/B, per-shape timings, ncu captures of pk_gemm.
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c1_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c1_bench_fused.json 2> $OUT/c1_bench_fused.err
RCOT_FUSED_GDFN_MID=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > $OUT/c1_bench_unfused.json 2> $OUT/c1_bench_unfused.err
timeout 300 python scratch/detail_prof.py 70 > $OUT/c1_detail.txt 2>&1
for a in "255 128" "127 128" "255 64" "510 32"; do timeout 120 python scratch/gf_one.py $a; done > $OUT/c1_gf.txt 2>&1
for a in "510 96 1" "288 96 1" "96 255 0" "96 96 0" "254 48 1"; do timeout 120 python scratch/pk_one.py $a; done > $OUT/c1_pk.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_gemm -s 3 -c 1 -o $OUT/pk_ln python scratch/pk_one.py 510 96 1 > $OUT/c1_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_gemm -s 3 -c 1 -o $OUT/pk_plain python scratch/pk_one.py 96 255 0 > $OUT/c1_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gdfn_mid -s 3 -c 1 -o $OUT/gf python scratch/gf_one.py 255 128 > $OUT/c1_ncu3.log 2>&1
for r in pk_ln pk_plain gf; do
  { python scripts/ncu_summarize.py $OUT/$r.ncu-rep --src 0 --top 40; python scripts/ncu_stalls.py $OUT/$r.ncu-rep 0; } > $OUT/c1_$r.txt 2>&1
done
ls -la $OUT
tail -5 $OUT/c1_tests.log
cat $OUT/c1_gf.txt $OUT/c1_pk.txt
