#!/bin/bash
# 8-GPU tuning of the block-cyclic Cholesky: grid x ring x NCCL channels; re-timed posterior (pipelined pushes)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29731"
run() { name=$1; shift; "$@" > gpurun_out/r02g_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/r02g_$name.log | cut -c1-420)"; }
run 2x4_ring2 $TR tools/dist_chol.py 131072 1024 --grid 2x4 --ring2 --reps 3
run 1x8_ring2_post $TR tools/dist_chol.py 131072 1024 --grid 1x8 --ring2 --reps 2 --post 10000
NCCL_MAX_NCHANNELS=4 run 2x4_ring2_nch4 $TR tools/dist_chol.py 131072 1024 --grid 2x4 --ring2 --reps 2
NCCL_MAX_NCHANNELS=8 run 2x4_ring2_nch8 $TR tools/dist_chol.py 131072 1024 --grid 2x4 --ring2 --reps 2
NCCL_MAX_NCHANNELS=4 run 1x8_ring2_nch4 $TR tools/dist_chol.py 131072 1024 --grid 1x8 --ring2 --reps 2
run 4x2_ring2 $TR tools/dist_chol.py 131072 1024 --grid 4x2 --ring2 --reps 2
run 2x4_nb2048 $TR tools/dist_chol.py 131072 2048 --grid 2x4 --ring2 --reps 2
run 2x4_nb512 $TR tools/dist_chol.py 131072 512 --grid 2x4 --ring2 --reps 2
