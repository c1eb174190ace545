#!/bin/bash
# round 2, GPU job N: table builder at 16 / 12 warps per SM; G2 bit-plane reduction kernels at 1 (ptxas) / 3 / 4 blocks per SM
O=gpurun_out/r02n; mkdir -p $O
P=snark_challenge_prover_reference_b200
for V in "" _red3 _red4; do
  B200_VERBOSE=1 B200_LIB=$P/libb200groth16$V.so timeout 600 python tools/time_query_msm.py 0 20 2 2 > $O/g2_reduce$V.jsonl 2> $O/g2_reduce$V.err
  B200_LIB=$P/libb200groth16$V.so timeout 600 python tools/profile_prove.py 0 20 > $O/prove4$V.log 2>&1
done
B200_LIB=$P/libb200groth16_red3.so timeout 600 python tools/profile_prove.py 1 15 > $O/prove6_red3.log 2>&1
for V in "" _red3 _red4; do echo "variant '$V'"; grep '"rep": 2' $O/g2_reduce$V.jsonl | cut -c1-200; grep "base table\|waited" $O/g2_reduce$V.err | cut -c1-120; grep " ms " $O/prove4$V.log | tail -n 1; done
grep " ms " $O/prove6_red3.log | tail -n 1
exit 0
