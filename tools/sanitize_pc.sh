#!/bin/bash
# compute-sanitizer memcheck over the fused PC kernel on a small workload (run under gpurun): tools/sanitize_pc.sh [molecules] [steps]
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/pc_time_probe.py ${1:-8} ${2:-2} > gpurun_out/sanitize_pc.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_pc.txt
grep -v "^=========     at\|Host Frame\|^=========         \|^=========$" gpurun_out/sanitize_pc.txt | head -60
