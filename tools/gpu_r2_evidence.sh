#!/bin/bash
# round-2 evidence pack: bench lines, A/B of the cross-attention kernels, ncu --set full of xattn_tc3, parity at the BASELINE
# shapes, memcheck of the kernels added this round
mkdir -p gpurun_out/ev
python bench.py > gpurun_out/ev/bench_r2_default.json 2> gpurun_out/ev/bench_r2_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/ev/bench_r2_reference.json 2>/dev/null
(for a in "1 100 529920" "4 100 529920" "1 100 132480" "4 100 132480" "1 100 33120" "4 100 33120" "36 100 14720" "36 200 14720" "144 100 3840" "144 100 960"; do timeout 300 python tools/prof_xattn_t.py $a; done; SPARSE=1 timeout 300 python tools/prof_xattn_t.py 1 100 529920; SPARSE=1 timeout 300 python tools/prof_xattn_t.py 4 100 529920) > gpurun_out/ev/xattn_t_ab_r2.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:xattn_tc3 -s 2 -c 1 -o gpurun_out/ev/ncu_r2_xattn_tc3 python tools/prof_xattn_t.py 4 100 529920 > /dev/null 2>&1
python -m pytest tests/test_decoder_gpu.py tests/test_brivis_pipeline_gpu.py -q -s -k "full_shape or cfg4 or cfg5b or cfg3" > gpurun_out/ev/parity_r2.txt 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_xattn_t_gpu.py tests/test_san_blocks_gpu.py -q -x -k "not 132480 and not 64000" > gpurun_out/ev/sanitizer_memcheck_r2.txt 2>&1
tail -3 gpurun_out/ev/sanitizer_memcheck_r2.txt; tail -3 gpurun_out/ev/parity_r2.txt; cat gpurun_out/ev/xattn_t_ab_r2.txt | cut -c1-150
