#!/bin/bash
# Rebuilds the committed profile summaries from the captures under gpurun_out/ (run in the build container after a gpurun call).
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/bench_default.txt profiles/r01_bench_default_stdout.txt
{
echo "# round 1 — launch list of \`python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-rgbd\` (timed region only)"
echo
echo "\`ncu --nvtx --nvtx-include \"timed/\" --metrics gpu__time_duration.sum --clock-control none --cache-control none\` (CSV: gpurun_out/r01_bench_launches.csv, not tracked).  The control step is one CUDA graph; ncu profiles its kernel nodes one by one and serialises them (the dynamics half of kin, \`kin_kernel<..., 2>\`, overlaps collide + manifest in a real run): compare SHARES."
echo
echo "| kernel | launches | total us | avg us | share |"
echo "|---|---|---|---|---|"
python tools/launch_shares.py gpurun_out/r01_bench_launches.csv 16 | tail -n +2
echo
echo "($(python tools/launch_shares.py gpurun_out/r01_bench_launches.csv 1 | head -1) in 20 timed control steps.)"
} > profiles/r01_bench_launch_list_summary.md
{
echo "# round 1 — pipelined substep kernels + rasteriser, PickCube-v1 4096 envs, steady random-action state"
echo
echo "One \`ncu --set full --import-source on --clock-control none\` capture per kernel (tools/prof_step.py / tools/prof_render.py; reports gpurun_out/prof_*.ncu-rep, not tracked).  Durations under \`--set full\` are cold-cache and replayed: compare with the warm launch list (r01_bench_launch_list_summary.md)."
for k in solve_kernel_r1c kin_kernel_r1b kin2_kernel_r1c collide_kernel_r1b manifest_kernel_r1b rowfill_kernel_r1c raster_r1c; do
  [ -f gpurun_out/prof_$k.ncu-rep ] || continue
  echo; echo "## $k"; echo; python tools/ncu_summary.py gpurun_out/prof_$k.ncu-rep
done
echo
echo "Reading: every kernel of the substep is latency bound at 4096 sub-scenes (issue-slot utilisation 7-37 %, DRAM traffic a few MB per launch, i.e. < 1 % of HBM bandwidth).  Phase B (solve) is a dependent chain per row visit (wait + short scoreboard = shuffles/shared memory); kin is a serial recursion per sub-scene; collide and rowfill wait on L2 (long scoreboard) with 17 % of the warp slots occupied.  The rasteriser is instruction bound (issue 60 %) with 30 % of its warp samples at the block barriers between its passes; it writes 746 MB per launch (12 B / pixel)."
} > profiles/r01_pipeline_kernels_ncu_summary.md
git rm -q --cached profiles/r01_raster_kernel_ncu_summary.md 2>/dev/null || true
echo refreshed
