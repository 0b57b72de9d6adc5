#!/bin/bash
# SM-budget / lanes sweep of the default workload (device-resident value only)
mkdir -p gpurun_out
: > gpurun_out/r2_budget.txt
run() {
  echo "== budget=$1 streams=$2 clips=$3" >> gpurun_out/r2_budget.txt
  OVIS_SM_BUDGET=$1 python bench.py --no-other-configs --no-cpu-baseline --no-e2e --steps 5 --streams $2 --clips $3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(round(d['value']), round(d['ms_per_step'],1))" >> gpurun_out/r2_budget.txt
}
run 148 2 4
run 140 2 4
run 120 2 4
run 100 2 4
run 74 2 4
run 140 3 4
run 100 3 4
run 140 4 2
run 140 2 2
run 140 2 8
cat gpurun_out/r2_budget.txt
