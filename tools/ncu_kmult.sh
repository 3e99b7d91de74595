#!/bin/bash
# ncu --set full capture of one k_mult update launch at 1024x1024x512 (source of roofline.traffic); the plain run comes first
set -e
timeout -s KILL 200 python tools/microbench.py c5 3 > gpurun_out/plain_kmult_c5.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_mult -s 4 -c 1 -o gpurun_out/prof_r02_kmult_c5 -f python tools/microbench.py c5 3 > gpurun_out/ncu_kmult_c5.log 2>&1
