#!/bin/bash
# BASELINE.json configs #3 (Cosine / Dice, n in {2,3,4}) and the skewed variant of #2 -> gpurun_out/<tag>/configs.jsonl
TAG=${1:-cfg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
: > $OUT/configs.jsonl
for args in "--metric Jaccard --ngram 3" "--metric Cosine --ngram 3" "--metric Dice --ngram 3" "--metric Jaccard --ngram 2" \
            "--metric Jaccard --ngram 4" "--metric Cosine --ngram 2" "--metric Cosine --ngram 4" "--metric Dice --ngram 2" \
            "--metric Dice --ngram 4" "--metric Jaccard --ngram 3 --data zipf" "--metric Cosine --ngram 3 --data zipf"; do
  line=$(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $args 2>>$OUT/configs.err)
  echo "$line" >> $OUT/configs.jsonl
  echo "$args => $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e6,2),"Mq/s  e2e",round(d["e2e"]["value"]/1e6,2),"Mq/s  alg KB/query",round(r["algorithmic_bytes_per_query"]/1e3,1)," GB/s",round(r["achieved"]), "frac",round(r["frac"],3),"matches",round(d["results"]["queries_with_a_match"],3))' 2>/dev/null)"
done
