#!/bin/bash
# Run under gpurun on the B200 box: plain run first, then the ncu launch list and one full capture of the sweep
# kernel (B200_PROFILING.md recipe).  Usage: tools/profile_gpu.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 1 --iterations 500 --burnin 100 --no-cpu --no-e2e $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:bvar_kernel\|tvp_tile\|bvar_kron -s 2 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
