#!/bin/bash
# sanitizer passes over the pipeline's tests, the two f-row bench lines, the whole GPU suite
O=gpurun_out/r2s; mkdir -p $O
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "words_dict or candidate_rows or cars_jaccard" > $O/racecheck.log 2>&1; tail -4 $O/racecheck.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "words_dict or candidate_rows" > $O/synccheck.log 2>&1; tail -4 $O/synccheck.log
timeout 600 python bench.py --workload spellchecker > $O/bench_spellchecker.json 2> $O/bench_spellchecker.err; echo "spell rc=$?"; cut -c1-400 $O/bench_spellchecker.json
timeout 600 python bench.py --workload autocomplete > $O/bench_autocomplete.json 2> $O/bench_autocomplete.err; echo "auto rc=$?"; cut -c1-400 $O/bench_autocomplete.json
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
