#!/bin/bash
# round 2, GPU job L: 4-lane G1 cooperative groups, cooperative reduction off inside concurrent proofs, external witness map
O=gpurun_out/r02l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 600 python tools/time_query_msm.py 1 15 3,2 0 > $O/mnt6_coop.jsonl 2> $O/mnt6_coop.err
timeout 600 python tools/profile_prove.py 1 15 > $O/prove6_coop.log 2>&1
timeout 600 python tools/profile_shard.py 0 20 7 4 > $O/shard_w7_coop.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 3 $O/pytest.log; tail -n 1 $O/mnt6_coop.jsonl; grep " ms " $O/prove6_coop.log | tail -n 1; grep " ms " $O/shard_w7_coop.log | tail -n 1;  head -c 250 $O/bench_n1.json
exit 0
