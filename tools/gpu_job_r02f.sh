#!/bin/bash
# round 2, GPU job F: full GPU suite (batch_exp, complete Groth16 proofs, shared L preparation), batch_exp throughput, bench line
O=gpurun_out/r02f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/time_batch_exp.py 20 > $O/batch_exp.jsonl 2> $O/batch_exp.err
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
B200_SHARE_PREP=0 timeout 600 python tools/profile_prove.py 0 20 > $O/prove_noshare.log 2>&1
timeout 600 python tools/profile_prove.py 0 20 > $O/prove_share.log 2>&1
tail -3 $O/pytest.log; cat $O/batch_exp.jsonl; head -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err; tail -3 $O/prove_share.log $O/prove_noshare.log
exit 0
