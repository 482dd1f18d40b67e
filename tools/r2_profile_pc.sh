#!/bin/bash
# Round-2 profile of the shipping fused PC kernel (run under gpurun, 1 GPU):
#   1. ncu --set full + source on a reduced workload (148 groups x 6 PC steps: one group per CTA)
#   2. DRAM traffic + duration of ONE launch at the bench configuration (1024 x 10 x 1000)
#   3. launch list of `bench.py --steps 2 --warmup 1 --skip-pretrain`
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sde2d3d_pc -s 1 -c 1 -f -o gpurun_out/r2_pc_full \
    python tools/pc_time_probe.py 148 6 > gpurun_out/r2_pc_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sde2d3d_pc -c 1 \
    --csv --log-file gpurun_out/r2_pc_traffic.csv python bench.py --steps 1 --warmup 0 --skip-pretrain --no-cpu-baseline > gpurun_out/r2_pc_traffic.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v11.csv \
    python bench.py --steps 2 --warmup 1 --skip-pretrain --no-cpu-baseline > gpurun_out/r2_launches_v11.log 2>&1
tail -3 gpurun_out/r2_pc_full.log; tail -5 gpurun_out/r2_pc_traffic.csv; tail -3 gpurun_out/r2_launches_v11.csv
