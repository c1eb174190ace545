#!/bin/bash
# round 2, GPU job T: GPU suite with the last test added
O=gpurun_out/r02t; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -n 5 $O/pytest.log
exit 0
