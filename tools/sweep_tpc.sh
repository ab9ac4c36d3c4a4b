#!/bin/bash
# lanes-per-chain sweep of a workload (run under gpurun): tools/sweep_tpc.sh <tag> <workload> "<tpc list>" [extra args]
TAG=$1; WL=$2; LIST=$3; shift 3
mkdir -p gpurun_out
for t in $LIST; do
  python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu --no-e2e --threads-per-chain $t "$@" > gpurun_out/tpc_${TAG}_$t.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tpc_${TAG}_$t.log").read().strip().splitlines()[-1])
    print("tpc=$t", d["roofline"]["kernel"], "draws/s=%.4g"%d["value"], "ms/step=%.3f"%d["ms_per_step"], "sweep_ms=%.3f"%d["roofline"]["sweep_kernel_ms_per_step"])
except Exception as e:
    print("tpc=$t failed", e); print(open("gpurun_out/tpc_${TAG}_$t.log").read()[-500:])
PY
done
