#!/bin/bash
# where the step goes on the heavy points of config #3: stage times + the pipeline's counters (SG_TRACE=1)
O=gpurun_out/r2h; mkdir -p $O
for a in "--data zipf --metric Cosine" "--data zipf --metric Jaccard" "--ngram 2 --metric Jaccard" "--ngram 2 --metric Cosine"; do
  echo "== $a"
  SG_TRACE=1 timeout 300 python tools/prof_step.py --calls 3 --stages $a > $O/out.txt 2>&1
  grep "sg_search_stage_times" $O/out.txt | tail -1 | cut -c1-200; grep "sg_tokens_count" $O/out.txt | cut -c1-330
  cat $O/out.txt >> $O/heavy_points.txt
done
