#!/bin/bash
# round 2, GPU job P: what a 19/64 slice of B2 (the bottleneck of the 8-GPU plan) spends where; cooperative reduction on / off,
# window width forced one up / down
O=gpurun_out/r02p; mkdir -p $O
S="0-0;0-0;0-19;0-0;0-0"
timeout 300 python tools/profile_spans.py "$S" > $O/b2_default.log 2>&1
B200_COOP=0 timeout 300 python tools/profile_spans.py "$S" > $O/b2_coop0.log 2>&1
B200_BATCH_AFFINE=0 timeout 300 python tools/profile_spans.py "$S" > $O/b2_xyzz.log 2>&1
for f in b2_default b2_coop0 b2_xyzz; do echo $f; grep -A1 " ms " $O/$f.log | tail -n 2 | cut -c1-250; done
exit 0
