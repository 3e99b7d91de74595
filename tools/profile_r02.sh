#!/bin/bash
# round-2 evidence on one B200: GPU test suite, default bench line, ncu launch list and --set full capture of one iteration at
# 512x512x256 (every command first runs plain, then under ncu)
set -x
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_final.log 2>&1; echo pytest rc=$?
( time timeout -s KILL 400 python bench.py ) > gpurun_out/bench_r02_final.log 2>&1; echo bench rc=$?
timeout -s KILL 120 python tools/microbench.py c4 4 > gpurun_out/plain_c4.log 2>&1 || exit 1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4_tma.csv python tools/microbench.py c4 4 > gpurun_out/ncu_launches.log 2>&1; echo launches rc=$?
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_dct_blu16|k_qstep|k_thomas|k_q2_fix|k_mult" -s 27 -c 9 -o gpurun_out/prof_r02_iter_c4_tma -f python tools/microbench.py c4 4 > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
tail -3 gpurun_out/pytest_r02_final.log; cat gpurun_out/plain_c4.log
