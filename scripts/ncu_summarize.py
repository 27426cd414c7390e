"""Summarise an .ncu-rep: key metrics per launch + the most-stalled SASS lines of launch `--src N`.
Usage: python scripts/ncu_summarize.py rep.ncu-rep [--src 1] [--top 30]"""
import argparse, csv, io, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--src", type=int, default=-1)
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
for li, r in enumerate(data):
    name = r[hdr.index('Kernel Name')][:80]
    print(f"---- launch {li}: {name}")
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"   {w} = {r[i]} {units[i]}")
if a.src >= 0:
    name = data[a.src][hdr.index('Kernel Name')]
    src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--launch-skip", str(a.src), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 5 and r[2].strip().isdigit()]
    seen, uniq = set(), []
    for r in rows:
        if r[0] in seen:
            continue
        seen.add(r[0]); uniq.append(r)
    tot = sum(int(r[2]) for r in uniq)
    print(f"==== top stalls, launch {a.src} ({len(uniq)} SASS instructions, {tot} samples)")
    for r in sorted(uniq, key=lambda r: -int(r[2]))[:a.top]:
        print(f"{int(r[2]):7d} {int(r[2]) / max(tot,1) * 100:5.1f}%  exec={r[5]:>9s}  {r[1].strip()[:90]}")
