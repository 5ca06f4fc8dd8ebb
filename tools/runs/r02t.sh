#!/bin/bash
timeout 300 python tools/splitk_ab.py > gpurun_out/r02t_splitk_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02t_splitk_ab.txt
timeout 300 python tools/chain_prof.py > gpurun_out/r02t_chain_prof.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/r02t_chain_prof.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_reference_goldens.py -m gpu -q -x > gpurun_out/r02t_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02t_parity.log
