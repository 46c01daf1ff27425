for args in "--metric Jaccard --ngram 3 --data zipf" "--metric Cosine --ngram 3 --data zipf" "--metric Jaccard --ngram 3"; do
  line=$(timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $args 2>/dev/null)
  echo "$args => $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e6,2),"Mq/s  e2e",round(d["e2e"]["value"]/1e6,2),"shift",d["config"]["bucket_shift"],r["stage_ms"])')"
done
