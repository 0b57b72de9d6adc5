#!/bin/bash
# PDL A/B: parity subset, then the default bench with and without programmatic dependent launch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_decoder_gpu.py tests/test_xattn_t_gpu.py tests/test_kernels_gpu.py tests/test_brivis_pipeline_gpu.py tests/test_temporal_gpu.py -m gpu -x -q > gpurun_out/pytest_pdl.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pdl.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; echo "bench pdl rc=$?"
OVIS_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; echo "bench nopdl rc=$?"
python - <<'PY'
import json
for n in ("pdl1", "pdl0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
        print(n, round(d["value"]), {k: v.get("ms_per_step") for k, v in d.get("kernels", {}).items()})
        for k, v in d.get("other_configs", {}).items():
            print("   ", k, round(v["value"]), v["kernels"].get("query_side"))
    except Exception as e:
        print(n, "failed", e)
PY
