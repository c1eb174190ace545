#!/bin/bash
# round 2, GPU job Q: thresholds after job P (XYZZ below 20 M entries for G2, thread-per-chunk reduction for G2 above 2^17
# buckets): the 19/64 and 40/64 slices of B2 again, the 40/64 slice in both accumulation modes; GPU suite; ncu of the new
# reduction kernel
O=gpurun_out/r02q; mkdir -p $O
timeout 300 python tools/profile_spans.py "0-0;0-0;0-19;0-0;0-0" > $O/b2_19.log 2>&1
timeout 300 python tools/profile_spans.py "0-0;0-0;0-40;0-0;0-0" > $O/b2_40.log 2>&1
B200_BATCH_AFFINE=0 timeout 300 python tools/profile_spans.py "0-0;0-0;0-40;0-0;0-0" > $O/b2_40_xyzz.log 2>&1
B200_BATCH_AFFINE=1 timeout 300 python tools/profile_spans.py "0-0;0-0;0-40;0-0;0-0" > $O/b2_40_affine.log 2>&1
for f in b2_19 b2_40 b2_40_xyzz b2_40_affine; do echo $f; grep -A1 " ms " $O/$f.log | tail -n 2 | cut -c1-250; done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -n 3 $O/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_rows_kernel -c 1 -o $O/reduce_rows_g1 \
  python tools/time_query_msm.py 0 20 3 0 > $O/ncu_rows.log 2>&1
ncu -i $O/reduce_rows_g1.ncu-rep --page raw --csv > $O/prof_reduce_rows_g1_r02_raw.csv 2>/dev/null
rm -f $O/reduce_rows_g1.ncu-rep
ls -la $O
exit 0
