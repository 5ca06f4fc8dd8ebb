#!/bin/bash
timeout 300 python tools/diag_time.py > gpurun_out/r02m_diag_time.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/r02m_diag_time.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02m_parity.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02m_parity.log
