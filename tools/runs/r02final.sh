#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02final_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02final_gpu_suite.log; tail -3 gpurun_out/r02final_gpu_suite.log
timeout 900 python bench.py > gpurun_out/r02final_bench_1gpu.json 2> gpurun_out/r02final_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02final_bench_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['stage_ms_per_step_serial'], d['op_path']['value'], json.dumps(d.get('single_eval_latency')), d['cpu_baseline'], d['gpu_launches'])"
timeout 300 python tools/latency2.py > gpurun_out/r02final_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02final_latency.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02final_launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-metric2 --no-extras > gpurun_out/r02final_ncu.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/r02final_launches_all.csv
