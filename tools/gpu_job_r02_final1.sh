#!/bin/bash
# round 2, final single-GPU evidence: ncu captures of the dominant kernels, launch list of one bench step, then the two
# bench arms back to back on the same box (reference first, like the driver does)
O=gpurun_out/r02_final; mkdir -p $O
nproc > $O/host.txt; lscpu | grep "Model name" >> $O/host.txt; nvidia-smi -L >> $O/host.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 1 -c 1 -o $O/prof_acc_g1 python tools/time_query_msm.py 0 20 3 0 > $O/ncu_g1.log 2>&1
ncu -i $O/prof_acc_g1.ncu-rep --page raw --csv > $O/prof_accumulate_g1_r02_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:msm_reduce_coop_kernel -c 1 -o $O/prof_coop python tools/time_query_msm.py 1 15 2 0 > $O/ncu_coop.log 2>&1
ncu -i $O/prof_coop.ncu-rep --page raw --csv > $O/prof_reduce_coop_fq3_r02_raw.csv 2>/dev/null
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
rm -f $O/prof_acc_g1.ncu-rep $O/prof_coop.ncu-rep
head -c 300 $O/bench_ref.json; echo; head -c 300 $O/bench_n1.json; echo; wc -l $O/launches_r02.csv
exit 0
