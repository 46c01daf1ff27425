#!/bin/bash
# GPU-box session for the direct result path: parity tests, the e2e sweep, one default bench line
TAG=${1:-direct}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
echo "== e2e sweep"; timeout 600 python tools/e2e_direct.py > $OUT/e2e_direct.txt 2> $OUT/e2e_direct.err; echo rc=$?; cat $OUT/e2e_direct.txt; tail -3 $OUT/e2e_direct.err
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?; cut -c1-200 $OUT/bench.json; tail -3 $OUT/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/direct/bench.json"))
print("value", d["value"] / 1e6, "e2e", d["e2e"], "cpu", d.get("cpu_baseline"))
P
