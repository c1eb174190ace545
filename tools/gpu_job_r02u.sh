#!/bin/bash
# round 2, GPU job U: ncu launch list of one bench step with the final build (per-launch times are cold-cache and serialised:
# the kernels' SHARES of the step are what it shows)
O=gpurun_out/r02u; mkdir -p $O
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r02_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
wc -l $O/launches_r02_final.csv
python tools/summarize_launches.py $O/launches_r02_final.csv > $O/launches_r02_final_summary.md 2>&1; head -n 25 $O/launches_r02_final_summary.md
exit 0
