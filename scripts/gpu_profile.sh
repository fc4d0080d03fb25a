#!/bin/bash
# Box-side profiling pass (run under gpurun): isolated kernel timings, the bench line, the ncu launch list of one eager step and the
# `--set full` metrics of the same step as CSV (no .ncu-rep: they exceed gpurun's 64 MiB return limit).   usage: gpu_profile.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python scripts/time_fused.py > gpurun_out/time_fused_$tag.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$tag.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python scripts/step_once.py 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --csv --page raw --log-file gpurun_out/full_$tag.csv python scripts/step_once.py 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
tail -45 gpurun_out/time_fused_$tag.log
