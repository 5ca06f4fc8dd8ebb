#!/bin/bash
# 8-GPU box: multi-rank correctness + exact GP N=131072 with the new diagonal-tile kernel / tile split / fused solves, then the bench
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu" > gpurun_out/r02x_multi_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_multi_gpu_tests.log; tail -3 gpurun_out/r02x_multi_gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29721"
timeout 400 $TR tools/dist_chol.py 131072 1024 --grid 2x4 --verify --reps 2 > gpurun_out/r02x_dist8_2x4.log 2>&1; tail -1 gpurun_out/r02x_dist8_2x4.log | cut -c1-700
timeout 400 $TR tools/dist_chol.py 131072 1024 --grid 1x8 --grad > gpurun_out/r02x_dist8_1x8_grad.log 2>&1; tail -1 gpurun_out/r02x_dist8_1x8_grad.log | cut -c1-900
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench8.json 2> gpurun_out/r02x_bench8.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02x_bench8.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], json.dumps(d.get('strong')), json.dumps(d.get('metric3'))[:900])"
