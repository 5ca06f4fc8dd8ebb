#!/bin/bash
timeout 120 ./tools/lat_bench 2>&1 | tail -1
timeout 300 python tools/diag_time.py > gpurun_out/r02p_diag_time.txt 2>&1; echo "diag rc=$?"; grep -v "variant 1\|B=   8\|B=  64" gpurun_out/r02p_diag_time.txt | head -8
timeout 1500 python bench.py > gpurun_out/r02p_bench_1gpu.json 2> gpurun_out/r02p_bench_1gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02p_bench_1gpu.json
