#!/bin/bash
# Round-2 ncu captures of the secondary kernels (run under gpurun, 1 GPU): SchNet CFConv (stress shapes), the TMA-fed GEMM,
# the fused dense inference kernels.  Summaries are extracted into profiles/ by tools/ncu_summary.py.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:schnet_cfconv -s 6 -c 1 -f -o gpurun_out/r2_cfconv python tools/stress_probe.py 512 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_tma -s 4 -c 1 -f -o gpurun_out/r2_tcgemm_tma python tools/tcperf.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"dense_attn_sym|dense_pair_mlp|dense_edge_final_mlp|dense_gcn" -s 40 -c 12 -f -o gpurun_out/r2_dense_fused python tools/dense_sampler_probe.py 256 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
