#!/bin/bash
timeout 300 python tools/chain_prof.py > gpurun_out/r02s_chain_prof.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/r02s_chain_prof.txt
timeout 400 python tools/sweep_block.py > gpurun_out/r02s_sweep_block.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02s_sweep_block.txt
timeout 400 python tools/sweep_batch.py > gpurun_out/r02s_sweep_batch.txt 2>&1; echo "sweepb rc=$?"; cat gpurun_out/r02s_sweep_batch.txt
