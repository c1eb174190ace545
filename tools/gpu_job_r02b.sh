#!/bin/bash
# round 2, GPU job B: GPU test-suite on the new batch-affine kernel + device bingcd, variant timings, ncu of round 1
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
P=$PWD/snark_challenge_prover_reference_b200
timeout 600 python tools/time_query_msm.py 0 20 3,2 0,1 > $O/variant.jsonl 2> $O/variant.err
B200_LIB=$P/libb200groth16_sqr.so timeout 600 python tools/time_query_msm.py 0 20 3,2 0,1 > $O/variant_sqr.jsonl 2> $O/variant_sqr.err
B200_LIB=$P/libb200groth16_aff4.so timeout 600 python tools/time_query_msm.py 0 20 3,2 1 > $O/variant_aff4.jsonl 2> $O/variant_aff4.err
timeout 600 python tools/time_query_msm.py 1 15 3,2 0,1 > $O/variant_mnt6.jsonl 2> $O/variant_mnt6.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_affine_round_kernel -c 2 -o $O/prof_affine_round_g1 python tools/time_query_msm.py 0 20 3 1 > $O/ncu.log 2>&1
ncu -i $O/prof_affine_round_g1.ncu-rep --page raw --csv > $O/prof_affine_round_g1_raw.csv 2>/dev/null
tail -3 $O/pytest.log; tail -n 2 $O/variant*.jsonl; tail -3 $O/ncu.log
exit 0
