#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
echo "== dbg"
SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_dbg.so SG_TRACE=1 timeout 300 python tools/repro_words.py 65536 > $O/repro_words_dbg.txt 2>&1; grep "equals" $O/repro_words_dbg.txt | cut -c1-400
echo "== default lib, lean forced"
timeout 300 python tools/repro_words.py 65536 > $O/repro_words_lean.txt 2>&1; grep "equals" $O/repro_words_lean.txt | cut -c1-400
timeout 300 python tools/repro_words.py 65536 > $O/repro_words_lean2.txt 2>&1; grep "equals" $O/repro_words_lean2.txt | cut -c1-400
