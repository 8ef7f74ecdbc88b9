"""Prints the key metrics + warp-state samples of one ncu report (first kernel): usage ncu_summary.py <rep> [title]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
o = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(o)))
h = r[0]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
print(f"kernel: {r[2][h.index('Kernel Name')]}\n\n| metric | unit | value |\n|---|---|---|")
for k in keys:
    if k in h:
        print(f"| {k} | {r[1][h.index(k)]} | {r[2][h.index(k)]} |")
st = [(float(r[2][i] or 0), k) for i, k in enumerate(h) if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('not_issued')]
t = sum(v for v, _ in st)
print("\nWarp-state samples:\n\n| state | share |\n|---|---|")
for v, k in sorted(st, reverse=True)[:7]:
    print(f"| {k[33:]} | {100 * v / t:.1f}% |")
