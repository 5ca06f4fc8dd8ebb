#!/bin/bash
timeout 300 python tools/latency2.py > gpurun_out/r02v_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02v_latency.txt
timeout 300 python tools/chain_prof.py > gpurun_out/r02v_chain_prof.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/r02v_chain_prof.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02v_gpu_suite.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms_per_step_serial'], d['op_path']['value'])"
