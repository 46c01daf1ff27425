#!/bin/bash
# bench every tuning build under suggest_b200/variants/ (SUGGEST_B200_LIB) next to the product library
mkdir -p gpurun_out/variants
for lib in "" suggest_b200/variants/*.so; do
  name=$(basename "${lib:-product}" .so)
  SUGGEST_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/variants/$name.json 2>/dev/null
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/variants/{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["value"] / 1e6, 1), "M q/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"] / 1e6, 1))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
