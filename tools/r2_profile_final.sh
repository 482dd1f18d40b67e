#!/bin/bash
# Final round-2 captures (run under gpurun, 1 GPU): launch lists of one pretraining step and of the 3D->2D PC steps, and an
# ncu --set full capture of the whole-chain pair-MLP training kernels.  Summaries -> profiles/ via tools/agg_launches.py / ncu_summary.py.
mkdir -p gpurun_out
PROBE_ONCE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_pretrain_launches_v15.csv python tools/pretrain_probe.py 256 1 > gpurun_out/r2_pretrain_launches_v15.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp3_train -s 10 -c 10 -f -o gpurun_out/r2_mlp3_train python tools/pretrain_probe.py 256 1 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file gpurun_out/r2_dense_launches_v15.csv python tools/dense_sampler_probe.py 256 20 > gpurun_out/r2_dense_launches_v15.log 2>&1
ls -la gpurun_out/*v15* gpurun_out/r2_mlp3_train.ncu-rep
