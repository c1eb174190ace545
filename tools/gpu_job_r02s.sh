#!/bin/bash
# round 2, GPU job S (final build): the north-star check on REAL challenge-size parameters, then the bench arms and the GPU suite.
#   ./generate_parameters (unmodified reference generator, full size: MNT4753 2^20, MNT6753 2^15)
#   ./main (unmodified reference prover)                      -> *-output-ref
#   b200_prove through the C ABI (Params.from_file + prove)    -> *-output-b200
#   the reference's own UNMODIFIED cuda_prover_piecewise.cu over the B:: bundle (oracle/_ref/piecewise_b200) -> *-output-piecewise
# sha256 of the three outputs per curve must be equal. Log and hashes -> gpurun_out/r02s/ (copied to profiles/).
ROOT=$PWD; O=$PWD/gpurun_out/r02s; mkdir -p $O
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
    if curve == 0:
        # the 8-GPU per-query plan of bench.py, every rank's share run on this one GPU, combined: must be the same proof
        import bench
        spans, _ = bench.query_plan(8)
        parts = b"".join(key.prove_partial_queries(inp, sp, bench.PLAN_UNITS, b1_scaled=True)[0] for sp in spans if sp)
        sharded = b.prove_combine(0, parts, sum(1 for sp in spans if sp), None)
        print("[b200 %s] 8-GPU per-query plan emulated on one GPU == unsharded proof: %s" % (name, sharded == proof))
        assert sharded == proof
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
cd $ROOT
rm -rf $W
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 1200 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
head -c 300 $O/bench_ref.json; echo; head -c 300 $O/bench_n1.json; echo; tail -n 3 $O/pytest.log
exit 0
