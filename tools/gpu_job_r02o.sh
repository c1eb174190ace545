#!/bin/bash
# round 2, GPU job O: per-query sharding (tests + single-GPU emulation of the 8- and 4-GPU plans), then job N's variants
O=gpurun_out/r02o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -n 3 $O/pytest.log
timeout 900 python tools/profile_plan.py 8 queries > $O/plan8_queries.log 2>&1; cat $O/plan8_queries.log | tail -n 12
timeout 900 python tools/profile_plan.py 4 queries > $O/plan4_queries.log 2>&1; cat $O/plan4_queries.log | tail -n 8
timeout 900 python tools/profile_plan.py 2 queries > $O/plan2_queries.log 2>&1; cat $O/plan2_queries.log | tail -n 5
bash tools/gpu_job_r02n.sh
exit 0
