#!/bin/bash
# round 2, GPU job E: full GPU suite on the current build, tail-fill A/B, Fq3 occupancy, full bench line with drop-in path
O=gpurun_out/r02e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/time_query_msm.py 0 20 3,2 0,1 > $O/variant.jsonl 2> $O/variant.err
B200_AFF_SPLIT=0 timeout 600 python tools/time_query_msm.py 0 20 3,2 1 > $O/variant_nosplit.jsonl 2> $O/variant_nosplit.err
timeout 600 python tools/time_query_msm.py 1 15 3,2 0 > $O/variant_mnt6.jsonl 2> $O/variant_mnt6.err
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -3 $O/pytest.log; tail -n 1 $O/variant*.jsonl; head -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err
exit 0
