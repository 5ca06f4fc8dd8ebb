#!/bin/bash
timeout 600 python -m pytest tests/test_theano_ops.py tests/test_theano_integration.py tests/test_gpu_parity.py -m gpu -q -x -k "ops or resume or integration or theano" > gpurun_out/r02spec_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02spec_tests.log
timeout 300 python tools/op_path_ab.py > gpurun_out/r02spec_op_path.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02spec_op_path.txt
