#!/bin/bash
# Profiling pass of one round (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are then copied
# into profiles/ by hand).  Follows /opt/skills/guides/B200_PROFILING.md: every ncu command runs only after the same
# command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-groth16 --no-strong"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
# 1. every launch of the bench command with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
echo "launch list rc=$?"
# 2. the dominant MSM kernel and the three NTT passes, full sets
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -f -o gpurun_out/prof_msm_acc_r2 \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-ntt --no-groth16 --no-strong > gpurun_out/prof_ncu2.log 2>&1
echo "msm_accumulate capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 9 -c 3 -f -o gpurun_out/prof_ntt_r2 \
    python tools/sweep.py ntt26 > gpurun_out/prof_ncu3.log 2>&1
echo "ntt capture rc=$?"
ls -la gpurun_out/*.ncu-rep
