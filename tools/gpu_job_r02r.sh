#!/bin/bash
# round 2, GPU job R: chunk length of the bit-plane reduction chosen by latency AND throughput; staggered start of the
# proofs of a batch (on / off, cooperative reduction on / off inside it)
O=gpurun_out/r02r; mkdir -p $O
timeout 600 python tools/time_query_msm.py 0 20 3,2 2 > $O/mnt4_reduce.jsonl 2> $O/mnt4_reduce.err
timeout 600 python tools/profile_prove.py 1 15 > $O/prove6.log 2>&1
timeout 600 python tools/profile_shard.py 0 20 7 4 > $O/shard_w7.log 2>&1
timeout 600 python tools/profile_prove.py 0 20 > $O/prove4.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_stagger.json 2> $O/bench_stagger.err
B200_BATCH_STAGGER=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_together.json 2> $O/bench_together.err
B200_COOP=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_stagger_coop0.json 2> $O/bench_stagger_coop0.err
grep '"rep": 2' $O/mnt4_reduce.jsonl | cut -c1-200
grep " ms " $O/prove6.log | tail -n 1; grep " ms " $O/shard_w7.log | tail -n 1; grep " ms " $O/prove4.log | tail -n 1
for f in bench_stagger bench_together bench_stagger_coop0; do echo $f; head -c 330 $O/$f.json | cut -c1-330; echo; tail -n 2 $O/$f.err; done
exit 0
