#!/bin/bash
# round 2: N-GPU step, default plan vs the witness map split over three ranks
N=$1; O=gpurun_out/r02_n${N}_b; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 2 > $O/bench.json 2> $O/bench.err
B200_BENCH_SPLIT_H=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 3 --warmup 2 > $O/bench_split_h.json 2> $O/bench_split_h.err
grep -o '"ms_per_step": [0-9.]*' $O/bench.json | head -1; grep -o '"ms_per_step": [0-9.]*' $O/bench_split_h.json | head -1; tail -n 3 $O/bench_split_h.err
exit 0
