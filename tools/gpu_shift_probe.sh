#!/bin/bash
# bucket-width probe: bench.py for one metric at forced widths
M=${1:-Cosine}
for s in 5 6 7 8; do
  line=$(SG_BUCKET_SHIFT=$s timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric $M 2>/dev/null)
  echo "$M shift $s => $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1),"Mq/s", d["roofline"]["stage_ms"])')"
done
