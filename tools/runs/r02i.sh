#!/bin/bash
# 1 GPU: compute-sanitizer racecheck on the small end-to-end case (plain run first)
python tools/sanitize_case.py > gpurun_out/r02i_plain.log 2>&1 && \
timeout 1100 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_case.py > gpurun_out/r02i_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02i_plain.log; grep -E "RACECHECK SUMMARY|SANITIZE_CASE_OK|Error|hazard" gpurun_out/r02i_racecheck.log | head -20; tail -5 gpurun_out/r02i_racecheck.log
