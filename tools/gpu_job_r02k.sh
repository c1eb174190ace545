#!/bin/bash
# round 2, GPU job J: lane-cooperative bucket reduction
O=gpurun_out/r02k; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 600 python tools/time_query_msm.py 1 15 3,2 0 > $O/mnt6_coop.jsonl 2> $O/mnt6_coop.err
B200_COOP=0 timeout 600 python tools/time_query_msm.py 1 15 3,2 0 > $O/mnt6_nocoop.jsonl 2> $O/mnt6_nocoop.err
B200_COOP=1 timeout 600 python tools/time_query_msm.py 0 20 3,2 2 > $O/mnt4_coop_forced.jsonl 2> $O/mnt4_coop_forced.err
timeout 600 python tools/profile_shard.py 0 20 7 4 > $O/shard_w7_coop.log 2>&1
B200_COOP=0 timeout 600 python tools/profile_shard.py 0 20 7 4 > $O/shard_w7_nocoop.log 2>&1
timeout 600 python tools/profile_prove.py 1 15 > $O/prove6_coop.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; B200_COOP=0 timeout 600 python tools/profile_prove.py 1 15 > $O/prove6_nocoop.log 2>&1
tail -n 3 $O/pytest.log; tail -n 1 $O/mnt6_coop.jsonl $O/mnt6_nocoop.jsonl $O/mnt4_coop_forced.jsonl; grep " ms " $O/shard_w7_coop.log | tail -n 1; grep " ms " $O/shard_w7_nocoop.log | tail -n 1; grep " ms " $O/prove6_coop.log | tail -n 1; head -c 250 $O/bench_n1.json
exit 0
