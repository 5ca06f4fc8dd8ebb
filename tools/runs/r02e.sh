#!/bin/bash
# 2-GPU box: the GPU suite (incl. 2-GPU block-cyclic checks), then the distributed Cholesky / gradient at size, then the bench
nvidia-smi -L
python -m pytest tests -m gpu -q > gpurun_out/r02e_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gpu_suite.log; tail -8 gpurun_out/r02e_gpu_suite.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29711"
$TR tools/dist_chol.py 65536 1024 --verify > gpurun_out/r02e_dist2_65536.log 2>&1; tail -1 gpurun_out/r02e_dist2_65536.log
$TR tools/dist_chol.py 65536 1024 --grid 2x1 > gpurun_out/r02e_dist2_65536_2x1.log 2>&1; tail -1 gpurun_out/r02e_dist2_65536_2x1.log
$TR tools/dist_chol.py 65536 1024 --grad > gpurun_out/r02e_dist2_65536_grad.log 2>&1; tail -1 gpurun_out/r02e_dist2_65536_grad.log
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --dist-n 65536 > gpurun_out/r02e_bench2.json 2> gpurun_out/r02e_bench2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02e_bench2.json
