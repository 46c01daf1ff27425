#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
echo "== classic"
SG_PIPELINE=classic SG_TRACE=1 timeout 300 python tools/repro_words.py 65536 > $O/repro_words_classic.txt 2>&1; grep "equals\|flagged\|sg_" $O/repro_words_classic.txt | grep -v "sg_search_batch:" | cut -c1-600
for v in bigflags bignodes; do
echo "== $v"
SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_$v.so SG_TRACE=1 timeout 300 python tools/repro_words.py 65536 > $O/repro_words_$v.txt 2>&1; grep "equals\|flagged\|sg_" $O/repro_words_$v.txt | grep -v "sg_search_batch:" | cut -c1-600
done
