#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 240 python -m pytest "tests/test_gpu_parity.py::test_candidate_rows_equal_separate_arrays" "tests/test_gpu_parity.py::test_chunked_arrival_equals_sliced_path" "tests/test_gpu_parity.py::test_pinned_result_buffers" -x -q > $O/pytest_rows.log 2>&1
echo "rows rc=$?"; tail -4 $O/pytest_rows.log; grep -B2 -A14 "Error" $O/pytest_rows.log | head -50
timeout 300 python tools/e2e_direct.py > $O/e2e_direct.txt 2>&1
grep Jaccard $O/e2e_direct.txt | head -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json'))
print('value %.1fM e2e %.1fM arrays %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6, d['e2e']['separate_id_and_score_arrays']['value']/1e6), d['roofline']['stage_ms'], d['results'])"
tail -3 $O/bench.err
