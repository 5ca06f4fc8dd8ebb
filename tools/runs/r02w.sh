#!/bin/bash
timeout 300 python tools/latency2.py > gpurun_out/r02w_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02w_latency.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_drivers.py -m gpu -q -x > gpurun_out/r02w_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02w_parity.log
timeout 900 python bench.py --no-cpu-baseline --no-metric2 > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms_per_step_serial'], d['op_path']['value'], d.get('config3'))"
