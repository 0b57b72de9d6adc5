#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_xattn_t_gpu.py -q -x > gpurun_out/r2_run3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.txt
tail -5 gpurun_out/r2_run3_pytest.txt
(for a in "1 100 529920" "4 100 529920" "1 100 132480" "4 100 33120" "36 100 14720" "144 100 3840"; do timeout 300 python tools/prof_xattn_t.py $a; done) > gpurun_out/r2_xattn_t_ab.txt 2>&1
cat gpurun_out/r2_xattn_t_ab.txt
