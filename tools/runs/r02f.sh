#!/bin/bash
# 8-GPU box: multi-rank correctness tests, exact GP N=131072 (2x4 and 1x8 grids, gradient, posterior), then the bench
nvidia-smi -L | wc -l
python -m pytest tests -m gpu -q -k "multi_gpu" > gpurun_out/r02f_multi_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_multi_gpu_tests.log; tail -4 gpurun_out/r02f_multi_gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29721"
$TR tools/dist_chol.py 131072 1024 --verify > gpurun_out/r02f_dist8_2x4.log 2>&1; tail -1 gpurun_out/r02f_dist8_2x4.log
$TR tools/dist_chol.py 131072 1024 --grid 1x8 --post 10000 --grad > gpurun_out/r02f_dist8_1x8_grad.log 2>&1; tail -1 gpurun_out/r02f_dist8_1x8_grad.log
$TR tools/dist_chol.py 131072 1024 --grid 1x8 --ring2 > gpurun_out/r02f_dist8_1x8_ring2.log 2>&1; tail -1 gpurun_out/r02f_dist8_1x8_ring2.log
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench8.json 2> gpurun_out/r02f_bench8.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench8.json')); print(d['value'], d['e2e']['value'], json.dumps(d.get('strong')), json.dumps(d.get('metric3')), json.dumps(d.get('config3')))"
