#!/bin/bash
# compute-sanitizer evidence for the flag-free, fence-free hand-over protocols of the sweep and
# factorisation kernels ("race detection").  One tool per GPU-box call
# (/opt/skills/guides/B200_PROFILING.md): tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
# Output: gpurun_out/sanitize_<tool>.log; copy the summary to profiles/ after reading it.
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 240 python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
OPMGPU_CLUSTER_COOP=0 timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
echo "exit $?" >> gpurun_out/sanitize_$tool.log
tail -25 gpurun_out/sanitize_$tool.log
