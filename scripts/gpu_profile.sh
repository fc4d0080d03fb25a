#!/bin/bash
# Box-side profiling pass (run under gpurun): isolated kernel timings, the bench lines of every config, the ncu launch list of one
# eager step and the `--set full` metrics of the same step as CSV (no .ncu-rep: they exceed gpurun's 64 MiB return limit).
#   usage: gpu_profile.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python scripts/time_fused.py > gpurun_out/time_fused_$tag.txt 2>&1
timeout 300 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_reference.json
for c in C3 C4 C5; do timeout 400 python bench.py --config $c --steps 20 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_$c.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_$tag.csv python scripts/step_once.py 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/net_launches_$tag.csv python scripts/net_once.py 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --csv --page raw --log-file gpurun_out/ncu_full_$tag.csv python scripts/step_once.py 1 > /dev/null 2>&1
timeout 300 python scripts/knn_sub_profile.py time > gpurun_out/knn_sub_$tag.txt 2>&1
ls -la gpurun_out | tail -12
python -c "
import json
for f in ['bench_$tag.json','bench_${tag}_C3.json','bench_${tag}_C4.json','bench_${tag}_C5.json','bench_${tag}_reference.json']:
    try:
        d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"
