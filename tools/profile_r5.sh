#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU).  Every command runs plain first and is profiled only if that run exits 0.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of the default bench command (shares of device time per kernel)
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r5_bench_plain.json 2> $O/r5_bench_plain.err &&
$NCU --metrics gpu__time_duration.sum -c 40000 --csv --log-file $O/r5_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r5_bench_ncu.log 2>&1
# (2) the wavefront kernels, full sections: kazen's WarmStudio.xml 512x512x64 and the 10^7-triangle scene 1920x1080x16
python tools/prof_paths.py warm > $O/r5_prof_warm_plain.log 2>&1 &&
KZ_PROF_LANES=1 $NCU --set full --import-source on --profile-from-start off -o $O/prof_r5_warm python tools/prof_paths.py warm > $O/r5_prof_warm_ncu.log 2>&1
python tools/prof_paths.py big 10000000 1920 1080 16 > $O/r5_prof_big_plain.log 2>&1 &&
KZ_PROF_LANES=1 $NCU --set full --import-source on --profile-from-start off -o $O/prof_r5_big python tools/prof_paths.py big 10000000 1920 1080 16 > $O/r5_prof_big_ncu.log 2>&1
# (3) k_trace on the 2^20 (SAH) and 10^7 (LBVH) soups at the bench's batch size
RES=4096 NINC=16777216 python tools/variant_bench.py > $O/r5_trace_1m_plain.log 2>&1 &&
RES=4096 NINC=16777216 $NCU --set full --import-source on -k regex:k_trace -s 4 -c 1 -o $O/prof_r5_trace_1m_primary python tools/variant_bench.py > /dev/null 2>&1
RES=4096 NINC=16777216 $NCU --set full --import-source on -k regex:k_trace -s 11 -c 1 -o $O/prof_r5_trace_1m_incoherent python tools/variant_bench.py > /dev/null 2>&1
TRIS=10000000 LBVH=1 RES=4096 NINC=16777216 python tools/variant_bench.py > $O/r5_trace_10m_plain.log 2>&1 &&
TRIS=10000000 LBVH=1 RES=4096 NINC=16777216 $NCU --set full --import-source on -k regex:k_trace -s 4 -c 1 -o $O/prof_r5_trace_10m_primary python tools/variant_bench.py > /dev/null 2>&1
TRIS=10000000 LBVH=1 RES=4096 NINC=16777216 $NCU --set full --import-source on -k regex:k_trace -s 11 -c 1 -o $O/prof_r5_trace_10m_incoherent python tools/variant_bench.py > /dev/null 2>&1
ls -la $O | tail -20
