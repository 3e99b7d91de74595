#!/bin/bash
# ncu --set full capture of k_mult (update launches) at 512x512x256; run only after the plain run has exited 0
set -e
python tools/microbench.py c4 4 > gpurun_out/plain_kmult.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mult -s 6 -c 2 -o gpurun_out/prof_r02_kmult_al -f python tools/microbench.py c4 4 > gpurun_out/ncu_kmult.log 2>&1
