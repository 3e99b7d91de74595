#!/bin/bash
# 2-GPU validation: NCCL parity (small grids incl. device weights, then configs[3] + the 3-level 256x256x128 solve), short bench at 512x512x256
T="timeout -s KILL"; R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T 150 $R --master-port 29601 tests/dist_parity.py > gpurun_out/dist2_r02final.log 2>&1; echo dist rc=$?; grep -a "dist parity\|Error" gpurun_out/dist2_r02final.log | head
$T 200 $R --master-port 29602 tests/dist_parity_big.py > gpurun_out/dist_big2_r02final.log 2>&1; echo big rc=$?; grep -a "dist parity\|Error" gpurun_out/dist_big2_r02final.log | head
if [ "$1" = "bench" ]; then
$T 120 $R --master-port 29603 bench.py --gpus 2 --workload c4 --no-cpu --no-e2e --no-ttt > gpurun_out/bench2_c4_r02final.log 2>&1; rc=$?; echo bench rc=$rc
grep -a -o "\"ms_per_step\": [0-9.]*\|\"poisson\": [0-9.]*\|\"k_mult\": [0-9.]*\|\"k_qstep\": [0-9.]*" gpurun_out/bench2_c4_r02final.log | tr "\n" " "; echo
fi
