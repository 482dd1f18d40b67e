#!/bin/bash
# compute-sanitizer memcheck over the training kernels + tcgen05 GEMM (run under gpurun)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_pretrain.py tests/test_gpu_tcgemm.py -m gpu -q -x -k "full or tcgemm or forward_linear or epilogue or backward_patterns" > gpurun_out/sanitize_train_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_train_memcheck.txt
tail -6 gpurun_out/sanitize_train_memcheck.txt
