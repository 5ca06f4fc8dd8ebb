#!/bin/bash
# 1 GPU: full suite, bench, ncu launch list of one timed step, ncu --set full of the two fast2 Gram kernels
python -m pytest tests -m gpu -q > gpurun_out/r02h_gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_gpu_suite.log; tail -6 gpurun_out/r02h_gpu_suite.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02h_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_per_step_serial'], d['gram_roofline']['frac'], d['gpu_launches'])"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-metric2 --no-extras"
$B > gpurun_out/r02h_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2760 -c 930 --csv --log-file gpurun_out/r02h_launches.csv $B > gpurun_out/r02h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_fwd_fast2 -s 4 -c 1 -o gpurun_out/r02h_prof_gram_fwd $B > gpurun_out/r02h_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_vjp_fast2 -s 4 -c 1 -o gpurun_out/r02h_prof_gram_vjp $B > gpurun_out/r02h_ncu3.log 2>&1
ls -la gpurun_out/ | grep r02h
