#!/bin/bash
# round 2: BASELINE.json config 5 on N GPUs (N = $1): sharded MSM sweep + NTT replicas
N=$1; O=gpurun_out/r02_sweep; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 2400 python tools/sweep_multi.py $2 > $O/sweep_n1.jsonl 2> $O/sweep_n1.err
else
  timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/sweep_multi.py $2 > $O/sweep_n$N.jsonl 2> $O/sweep_n$N.err
fi
grep -c '"kind"' $O/sweep_n$N.jsonl; tail -2 $O/sweep_n$N.jsonl; tail -3 $O/sweep_n$N.err
exit 0
