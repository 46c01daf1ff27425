#!/bin/bash
# bucket width against the pipeline on the heavy points of config #3 (Zipf-lettered text, 2-gram index)
O=gpurun_out/r2w; mkdir -p $O; : > $O/shift_heavy.txt
for a in "--data zipf --metric Jaccard" "--data zipf --metric Cosine" "--ngram 2 --metric Jaccard" "--ngram 2 --metric Cosine" "--ngram 4 --metric Cosine"; do
  for bs in 2 3 4 5 6 7; do
    SG_BUCKET_SHIFT=$bs SG_TRACE=1 timeout 300 python tools/prof_step.py --calls 3 --stages $a > $O/out.txt 2>&1
    echo "$a shift $bs: $(grep sg_search_stage_times $O/out.txt | tail -1 | sed 's/.*queries: //' | cut -c1-90) | $(grep sg_tokens_count $O/out.txt | sed "s/.*'sg_tokens_count_kernel': \([0-9.]*\), 'sg_resolve_kernel': \([0-9.]*\), 'sg_bitmap_search_kernel': \([0-9.]*\)}.*/count \1 resolve \2 fallback \3/")" | tee -a $O/shift_heavy.txt
  done
done
