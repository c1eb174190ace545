#!/bin/bash
# round 2, GPU job V: multi-threaded file reads in the loaders - the driver tests (B::read_params / read_input from files,
# golden outputs) and the drop-in path at full size
O=gpurun_out/r02v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -x -q > $O/pytest_driver.log 2>&1; echo "rc=$?" >> $O/pytest_driver.log
tail -n 2 $O/pytest_driver.log
timeout 600 python - > $O/drop_in.json 2> $O/drop_in.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
files = bench.ensure_synth(20, 15)
print(json.dumps(bench.drop_in_run(files)))
PY
cat $O/drop_in.json; tail -n 3 $O/drop_in.err
exit 0
