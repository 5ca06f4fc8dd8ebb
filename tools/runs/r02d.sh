#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r02d_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_gpu_suite.log; tail -12 gpurun_out/r02d_gpu_suite.log
python tools/dist_chol.py 4096 256 --check --verify --reps 1 > gpurun_out/r02d_dist1_check.log 2>&1; tail -3 gpurun_out/r02d_dist1_check.log
python tools/dist_chol.py 32768 1024 --grad --reps 1 > gpurun_out/r02d_dist1_grad.log 2>&1; tail -1 gpurun_out/r02d_dist1_grad.log
