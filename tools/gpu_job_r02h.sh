#!/bin/bash
# round 2, GPU job H: full GPU suite + bench line on the build with the Jacobian table builder, the high-priority
# preparation stream and the warm-up proof in B::read_params
O=gpurun_out/r02h; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python tools/profile_prove.py 0 20 > $O/prove.log 2>&1
timeout 600 python tools/profile_prove.py 1 15 > $O/prove6.log 2>&1
tail -3 $O/pytest.log; head -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err; tail -4 $O/prove.log $O/prove6.log
exit 0
