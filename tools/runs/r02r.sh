#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02r_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_gpu_suite.log; tail -4 gpurun_out/r02r_gpu_suite.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_diag2 -s 3 -c 1 -o gpurun_out/r02r_diag2 python tools/diag_time.py > gpurun_out/r02r_ncu_diag2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02r_ncu_diag2.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
