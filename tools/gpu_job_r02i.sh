#!/bin/bash
# round 2, GPU job I: narrower windows (recalibrated model), H-first issue order emulated per rank on one GPU
O=gpurun_out/r02i; mkdir -p $O
B200_H_FIRST=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prover or table or concurrent" > $O/pytest_hfirst.log 2>&1; echo "rc=$?" >> $O/pytest_hfirst.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or prover or table or merged" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for w in 7 3; do
  timeout 600 python tools/profile_shard.py 0 20 $w 4 > $O/shard_w$w.log 2>&1
  B200_H_FIRST=1 timeout 600 python tools/profile_shard.py 0 20 $w 4 > $O/shard_w${w}_hfirst.log 2>&1
done
timeout 600 python tools/time_query_msm.py 0 20 3,2 2 > $O/variant.jsonl 2> $O/variant.err
B200_VERBOSE=1 timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -2 $O/pytest_hfirst.log $O/pytest.log; grep " ms " $O/shard_w7.log | tail -2; grep " ms " $O/shard_w7_hfirst.log | tail -2; tail -n 2 $O/variant.jsonl; head -c 250 $O/bench_n1.json
exit 0
