#!/bin/bash
# final round-2 evidence pack (final build): bench lines (default / reference arm), launch list + DRAM traffic of one cfg-2 clip,
# ncu --set full of the dominant family's kernel (maskfeat_prep_tma_kernel) and of the wide chain, parity at the BASELINE
# shapes, memcheck over this round's new kernels
mkdir -p gpurun_out/fin
python bench.py > gpurun_out/fin/bench_r2_final.json 2> gpurun_out/fin/bench_r2_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin/bench_r2_final_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/fin/launches_r2_final_cfg2.csv \
  python bench.py --workload openvis_video_36x720x1280_q100_k40 --clips 1 --streams 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/fin/launches_r2_final_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:maskfeat_prep_tma -s 1 -c 1 -o gpurun_out/fin/ncu_r2_maskfeat_prep python tools/prof_gemm.py prep > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_chain_wide -s 12 -c 1 -o gpurun_out/fin/ncu_r2_chain_wide python tools/prof_chain.py > /dev/null 2>&1
python -m pytest tests/test_decoder_gpu.py tests/test_brivis_pipeline_gpu.py tests/test_pixel_decoder_gpu.py -q -s -k "full_shape or cfg4 or cfg5b or cfg3 or pixel_decoder" > gpurun_out/fin/parity_r2_final.txt 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_pixel_decoder_gpu.py tests/test_msda_gpu.py tests/test_postprocess_gpu.py tests/test_decoder_gpu.py -q -x -k "not full_shape and not cfg4 and not cfg5b and not real_shape and not 1025" > gpurun_out/fin/sanitizer_memcheck_r2_final.txt 2>&1
tail -3 gpurun_out/fin/sanitizer_memcheck_r2_final.txt; tail -3 gpurun_out/fin/parity_r2_final.txt; ls -la gpurun_out/fin
