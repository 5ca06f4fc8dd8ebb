#!/bin/bash
timeout 300 python tools/diag_time.py > gpurun_out/r02o_diag_time.txt 2>&1; echo "diag rc=$?"; grep -v "variant 1\|B=   8\|B=  64" gpurun_out/r02o_diag_time.txt | head -8
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02o_gpu_suite.log
timeout 300 python tools/latency2.py > gpurun_out/r02o_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02o_latency.txt
