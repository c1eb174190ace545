#!/bin/bash
# round 2, GPU job A: GPU test-suite, same-config reference arm (full size), b200 arm, kernel-variant timings
O=gpurun_out/r02a; mkdir -p $O
nproc > $O/host.txt; lscpu | grep "Model name" >> $O/host.txt; nvidia-smi -L >> $O/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
P=snark_challenge_prover_reference_b200
for v in "" _sqr _bingcd _sqrbingcd; do
  B200_LIB=$PWD/$P/libb200groth16$v.so timeout 600 python tools/time_query_msm.py 0 20 3,2 0,1 > $O/variant$v.jsonl 2> $O/variant$v.err
done
tail -3 $O/pytest.log; head -c 600 $O/bench_ref.json; echo; head -c 400 $O/bench_n1.json; echo; tail -2 $O/variant*.jsonl
