#!/bin/bash
timeout 600 python tools/oz_time.py > gpurun_out/r02k_oz_time.log 2>&1; echo "rc=$?"; cat gpurun_out/r02k_oz_time.log
