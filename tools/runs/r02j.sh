#!/bin/bash
timeout 600 python tools/oz_check.py --time > gpurun_out/r02j_oz_check.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02j_oz_check.log
