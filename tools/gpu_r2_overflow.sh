#!/bin/bash
# the scratch-overflow path of the count -> resolve pipeline at full size: words.dict, pipeline forced, 65,536 queries, every
# host / device result path against the oracle on all queries (tools/repro_words.py), twice
O=gpurun_out/r2o; mkdir -p $O
for i in 1 2; do
  timeout 300 python tools/repro_words.py 65536 > $O/repro_words_lean$i.txt 2>&1; grep "equals" $O/repro_words_lean$i.txt | cut -c1-400
done
