#!/bin/bash
# round 2, GPU job D: G1 batch-affine round kernel with shared-memory temporaries
O=gpurun_out/r02d; mkdir -p $O
P=$PWD/snark_challenge_prover_reference_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or prover or table or merged" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/time_query_msm.py 0 20 3,2 0,1 > $O/variant.jsonl 2> $O/variant.err
B200_LIB=$P/libb200groth16_g1b3.so timeout 600 python tools/time_query_msm.py 0 20 3 1 > $O/variant_g1b3.jsonl 2> $O/variant_g1b3.err
timeout 600 python tools/time_query_msm.py 1 15 3 0,1 > $O/variant_mnt6.jsonl 2> $O/variant_mnt6.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_affine_round_g1_kernel -c 1 -o $O/prof_affine_round_g1 python tools/time_query_msm.py 0 20 3 1 > $O/ncu.log 2>&1
ncu -i $O/prof_affine_round_g1.ncu-rep --page raw --csv > $O/prof_affine_round_g1_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_affine_round_kernel -c 1 -o $O/prof_affine_round_g2 python tools/time_query_msm.py 0 20 2 1 > $O/ncu2.log 2>&1
ncu -i $O/prof_affine_round_g2.ncu-rep --page raw --csv > $O/prof_affine_round_g2_raw.csv 2>/dev/null
tail -3 $O/pytest.log; tail -n 2 $O/variant*.jsonl
exit 0
