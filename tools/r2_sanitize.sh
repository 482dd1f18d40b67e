#!/bin/bash
# Round-2 compute-sanitizer pass over the kernels added this round (run under gpurun): memcheck on the dense fused / tcgen05 head /
# node-side kernels, the tcgen05 CFConv, the TMA-fed GEMM (small shapes), the step-wise sampler updates and the double-backward tape.
mkdir -p gpurun_out
run() {  # name, test path, -k expression
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest -x -q "$2" -k "$3" > gpurun_out/r2_sanitize_$1.txt 2>&1
  echo "memcheck rc=$?" >> gpurun_out/r2_sanitize_$1.txt
  grep -E "passed|failed|ERROR SUMMARY|memcheck rc" gpurun_out/r2_sanitize_$1.txt | tail -4
}
run dense tests/test_gpu_dense.py "dense"
run schnet tests/test_gpu_schnet.py "schnet"
run tcgemm tests/test_gpu_tcgemm.py "epilogue_and_strided or 257-65-64 or 1000-200-128 or 130-65-33 or 777-119-728"
run stepwise tests/test_gpu_sde2d3d.py "stepwise or large_group"
run force tests/test_variants.py "energy_force and mse"
