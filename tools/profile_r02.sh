#!/bin/bash
# round-2 evidence on one B200: GPU test suite and the default bench line (the ncu launch list / --set full captures of
# profiles/r02_* were taken by the earlier version of this script: plain run of tools/microbench.py c4 4 first, then the same
# command under ncu --metrics gpu__time_duration.sum and under ncu --set full; tools/ncu_kmult.sh for the 1024x1024x512 launch)
set -x
timeout -s KILL 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_final.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_r02_final.log
( time timeout -s KILL 400 python bench.py ) > gpurun_out/bench_r02_final.log 2>&1; echo bench rc=$?
