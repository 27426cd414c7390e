#!/bin/bash
# Round profile evidence, run on the GPU box through gpurun:  bash scripts/make_profiles.sh r1
# Produces (in gpurun_out/, copied to profiles/ afterwards):
#   launches_<r>_final.csv        every launch of ONE iteration with gpu__time_duration (B=32, P=128)
#   step_<r>_final_summary.txt    per-kernel totals / shares of that list
#   pm_gemm_<r>_final_summary.txt ncu --set full key metrics + stall breakdown of the level-1 pm_gemm launches
#   traffic_<r>.json              DRAM bytes per launch of the dominant kernel (pm_gemm) over one iteration
set -u
R=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
N=$(python scripts/profile_step.py --warm 0 | tail -1 | awk '{print $NF}')   # launches per iteration (incl. memsets)
echo "launches per iteration (host count): $N"
# 1) launch list of the 2nd iteration (kernels only; memsets are not kernels, so skip by kernel count of iteration 1)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/all_$R.csv python scripts/profile_step.py --warm 1 > $OUT/p1.log 2>&1
python - "$OUT/all_$R.csv" "$OUT/launches_${R}_final.csv" "$OUT/step_${R}_final_summary.txt" <<'PY'
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
half = len(data) // 2                      # two identical iterations were run: keep the second (warm) one
data = data[half:]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f); w.writerow(hdr); w.writerows(data)
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[kn]))
    v = float(r[mv].replace(',', '')); v = v / 1e6 if r[mu] == 'ns' else v / 1e3 if r[mu] == 'us' else v
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(sys.argv[3], 'w') as f:
    f.write(f"one adversarial iteration, B=32, P=128, ncu gpu__time_duration.sum (serialised, cold-ish caches)\n")
    f.write(f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} kernel launches\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1]:10.2f} ms {v[1] / tot * 100:5.1f}% {v[0]:5d}  {k[:120]}\n")
print(open(sys.argv[3]).read()[:1500])
PY
rm -f $OUT/all_$R.csv
# 2) full-set capture of the first level-1 pm_gemm launches of the 2nd iteration's forward
NPM=$(grep -c pm_gemm_kernel $OUT/launches_${R}_final.csv)
ncu --set full --clock-control none --import-source on -k regex:pm_gemm_kernel -s $((NPM + 1)) -c 5 -o $OUT/pm_final python scripts/profile_step.py --warm 1 > $OUT/p2.log 2>&1
{ python scripts/ncu_summarize.py $OUT/pm_final.ncu-rep --src 1 --top 25; for i in 0 1 2 3 4; do python scripts/ncu_stalls.py $OUT/pm_final.ncu-rep $i; done; } > $OUT/pm_gemm_${R}_final_summary.txt 2>&1
rm -f $OUT/pm_final.ncu-rep
# 3) DRAM traffic of every pm_gemm launch of one iteration
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pm_gemm_kernel -s $NPM -c $NPM --csv --log-file $OUT/pm_traffic.csv python scripts/profile_step.py --warm 1 > $OUT/p3.log 2>&1
python - "$OUT/pm_traffic.csv" "$OUT/traffic_$R.json" <<'PY'
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
mn, mv, mu = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot = sum(float(r[mv].replace(',', '')) * scale.get(r[mu], 1) for r in data)
n = len({r[0] for r in data})
json.dump({"kernel": "pm_gemm", "per_gpu_batch": 32, "patch": 128, "launches": n, "dram_bytes_total": tot,
           "dram_bytes_per_launch": tot / max(n, 1),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every pm_gemm launch of one iteration"},
          open(sys.argv[2], 'w'), indent=1)
print(open(sys.argv[2]).read())
PY
rm -f $OUT/pm_traffic.csv
ls -la $OUT
