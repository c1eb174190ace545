#!/bin/bash
# round 2, GPU job M: bit-plane bucket reduction (on / off), r * Bt1 in B1's tail, scaled partials
O=gpurun_out/r02m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for P in 1 0; do
  B200_REDUCE_PLANES=$P timeout 600 python tools/time_query_msm.py 0 20 3,2 0,2 > $O/mnt4_planes$P.jsonl 2> $O/mnt4_planes$P.err
  B200_REDUCE_PLANES=$P timeout 600 python tools/profile_shard.py 0 20 7 4 > $O/shard_w7_planes$P.log 2>&1
  B200_REDUCE_PLANES=$P timeout 600 python tools/profile_prove.py 1 15 > $O/prove6_planes$P.log 2>&1
done
timeout 600 python tools/profile_prove.py 0 20 > $O/prove4.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 3 $O/pytest.log
for P in 1 0; do echo "planes=$P"; grep '"rep": 2' $O/mnt4_planes$P.jsonl | cut -c1-220; grep " ms " $O/shard_w7_planes$P.log | tail -n 1; grep " ms " $O/prove6_planes$P.log | tail -n 1; done
grep " ms " $O/prove4.log | tail -n 1
head -c 250 $O/bench_n1.json
exit 0
