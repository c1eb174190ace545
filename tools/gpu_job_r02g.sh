#!/bin/bash
# round 2, GPU job G: the north-star check on REAL challenge-size parameters.
#   ./generate_parameters (unmodified reference generator, full size: MNT4753 2^20, MNT6753 2^15)
#   ./main (unmodified reference prover)                      -> *-output-ref
#   b200_prove through the C ABI (Params.from_file + prove)    -> *-output-b200
#   the reference's own UNMODIFIED cuda_prover_piecewise.cu over the B:: bundle (oracle/_ref/piecewise_b200) -> *-output-piecewise
# sha256 of the three outputs per curve must be equal. Log and hashes -> gpurun_out/r02g/ (copied to profiles/).
O=$PWD/gpurun_out/r02g; mkdir -p $O
REF=$PWD/oracle/_ref
W=/tmp/realparams; mkdir -p $W; cd $W
{
echo "host: $(nproc) cores, $(lscpu | grep 'Model name' | sed 's/.*: *//')"; nvidia-smi -L
t0=$(date +%s)
OMP_NUM_THREADS=$(nproc) $REF/generate_parameters > gen.log 2>&1
echo "generate_parameters (full): $(( $(date +%s) - t0 )) s"; ls -l MNT*-parameters MNT*-input
for c in MNT4753 MNT6753; do
  OMP_NUM_THREADS=$(nproc) $REF/main $c compute $c-parameters $c-input $c-output-ref > main_$c.log 2>&1
  grep -E "load params|Total time from input to output|multiexp +\[|polynomial H +\[" main_$c.log | sed "s/^/[main $c] /"
done
} > $O/real_params.log 2>&1
cd - > /dev/null
python - >> $O/real_params.log 2>&1 <<'PY'
import hashlib, os, time, json, sys
sys.path.insert(0, os.getcwd())
import snark_challenge_prover_reference_b200 as b
b.check(b.lib().b200_set_device(0))
W = "/tmp/realparams"
for curve, name in enumerate(("MNT4753", "MNT6753")):
    t0 = time.time(); key = b.Params.from_file(curve, os.path.join(W, name + "-parameters")); t1 = time.time()
    pre = key.precompute(0, 1)
    inp = open(os.path.join(W, name + "-input"), "rb").read()
    key.prove(inp)
    t2 = time.time(); proof, tm = key.prove(inp, timings=True); t3 = time.time()
    open(os.path.join(W, name + "-output-b200"), "wb").write(proof)
    print("[b200 %s] load params %.0f ms (%s), base tables %.1f s, prove %.1f ms" % (name, 1e3 * (t1 - t0), json.dumps({k: round(v) for k, v in key.load_ms().items()}), pre, 1e3 * (t3 - t2)))
    key.close()
PY
cd $W
for c in MNT4753 MNT6753; do
  B200_BUNDLE_TIMING=1 $REF/piecewise_b200 $c compute $c-parameters $c-input $c-output-piecewise 2>&1 | grep -E "load params|Total time" | sed "s/^/[piecewise_b200 $c] /" >> $O/real_params.log
done
sha256sum MNT*-output-* | tee $O/real_params.sha256 >> $O/real_params.log
for c in MNT4753 MNT6753; do
  n=$(sha256sum $c-output-ref $c-output-b200 $c-output-piecewise | awk '{print $1}' | sort -u | wc -l)
  echo "$c: $n distinct sha256 among reference main / b200_prove / piecewise over the B:: bundle (1 = bit-identical)" >> $O/real_params.log
done
cat $O/real_params.log
exit 0
