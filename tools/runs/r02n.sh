#!/bin/bash
timeout 300 python tools/diag_time.py > gpurun_out/r02n_diag_time.txt 2>&1; echo "diag rc=$?"; grep -v "variant 1" gpurun_out/r02n_diag_time.txt
