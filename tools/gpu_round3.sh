#!/bin/bash
# One gpurun call: default bench, the HBM-resident 10^7-triangle bench, then (only after those exited 0) ncu passes of the same commands.
set -u
O=gpurun_out
python bench.py > $O/bench10.json 2> $O/bench10.err; echo "bench rc=$?"
python bench.py --tris 10000000 --builder lbvh --no-paths --steps 5 > $O/bench_10m_r3.json 2> $O/bench_10m_r3.err; echo "bench10m rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench_r3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_r3_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 6 -c 2 -o $O/prof_trace_10m_r3 -f \
    python bench.py --tris 10000000 --builder lbvh --no-paths --no-cpu --steps 2 --warmup 3 > $O/ncu_r3_full10m.log 2>&1; echo "ncu full 10m rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 6 -c 2 -o $O/prof_trace_r3 -f \
    python bench.py --no-paths --no-cpu --steps 2 --warmup 3 > $O/ncu_r3_full.log 2>&1; echo "ncu full rc=$?"
tail -c 300 $O/bench10.err $O/bench_10m_r3.err
