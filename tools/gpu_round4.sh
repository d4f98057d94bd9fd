#!/bin/bash
# ncu passes of the tree as committed (run only after the same commands exited 0 without ncu): launch list of the bench, full capture of k_trace.
set -u
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > $O/plain_r4.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench_r4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_r4_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 6 -c 2 -o $O/prof_trace_r4 -f \
    python bench.py --no-paths --no-cpu --steps 2 --warmup 3 > $O/ncu_r4_full.log 2>&1; echo "ncu full rc=$?"
