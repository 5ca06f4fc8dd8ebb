#!/bin/bash
# 2-GPU box: multi-rank checks with the new diagonal-tile kernel on the panel chain, N=65536 verify + gradient
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu" > gpurun_out/r02q_multi_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_multi_gpu_tests.log; tail -3 gpurun_out/r02q_multi_gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29711"
timeout 600 $TR tools/dist_chol.py 65536 1024 --verify > gpurun_out/r02q_dist2_65536.log 2>&1; tail -1 gpurun_out/r02q_dist2_65536.log
timeout 600 $TR tools/dist_chol.py 65536 1024 --grad > gpurun_out/r02q_dist2_65536_grad.log 2>&1; tail -1 gpurun_out/r02q_dist2_65536_grad.log
