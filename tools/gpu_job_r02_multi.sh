#!/bin/bash
# round 2: the step on N GPUs of one box (N = $1), one process per GPU over NCCL
N=$1; O=gpurun_out/r02_n$N; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
if [ -n "$2" ]; then
  B200_BENCH_MNT6_MODE=$2 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_other_mode.json 2> $O/bench_other_mode.err
fi
head -c 400 $O/bench.json; echo; tail -3 $O/bench.err; head -c 300 $O/bench_other_mode.json 2>/dev/null
exit 0
