#!/bin/bash
# one-GPU check: GPU suite, bench, single-GPU block-cyclic Cholesky
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_gpu_suite.log; tail -5 gpurun_out/r02c_gpu_suite.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_per_step_serial'], d.get('op_path'), d.get('config3'), d['metric2']['value'])"
python tools/dist_chol.py 32768 1024 --verify > gpurun_out/r02c_dist1_32768.log 2>&1; tail -2 gpurun_out/r02c_dist1_32768.log
python tools/dist_chol.py 65536 1024 --reps 1 > gpurun_out/r02c_dist1_65536.log 2>&1; tail -1 gpurun_out/r02c_dist1_65536.log
