#!/bin/bash
python tools/latency2.py > gpurun_out/r02l_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02l_latency.txt
python -m pytest tests -m gpu -q -x > gpurun_out/r02l_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_gpu_suite.log; tail -6 gpurun_out/r02l_gpu_suite.log
