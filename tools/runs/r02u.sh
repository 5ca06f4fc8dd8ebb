#!/bin/bash
timeout 400 python tools/tile_split_ab.py > gpurun_out/r02u_tile_split_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02u_tile_split_ab.txt
timeout 300 python tools/chain_prof.py > gpurun_out/r02u_chain_prof.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/r02u_chain_prof.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02u_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_parity.log
