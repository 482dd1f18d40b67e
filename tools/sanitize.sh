#!/bin/bash
# compute-sanitizer pass over the main kernels (run under gpurun): memcheck on the smoke path + a small dense/SchNet run.
set -o pipefail
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.txt
tail -5 gpurun_out/sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.txt
tail -5 gpurun_out/sanitize_racecheck.txt
