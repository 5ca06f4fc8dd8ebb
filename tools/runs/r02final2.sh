#!/bin/bash
timeout 300 python tools/latency2.py > gpurun_out/r02final2_latency.txt 2>&1; echo "latency rc=$?"; cat gpurun_out/r02final2_latency.txt
timeout 600 python -m pytest tests/test_widened_features.py tests/test_reference_goldens.py tests/test_gpu_drivers.py -m gpu -q -x > gpurun_out/r02final2_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02final2_tests.log
