#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): tools/profile.sh launches|paths|trace
# Every command runs plain first and is profiled only if that run exits 0.  Reports are condensed on the box (tools/ncu_summary.py) and
# only the ones needed for source-level views are kept, so that gpurun_out/ stays far below its 64 MiB limit.
set -u
O=gpurun_out
R=${R:-r6}
NCU="ncu --clock-control none"
case "${1:-}" in
launches)
  # one steady-state step of the headline at spec (KZ_PROFILE_STEP brackets the timed step with cudaProfilerStart/Stop)
  python bench.py --steps 1 --warmup 3 --no-cpu --legs headline > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err &&
  KZ_PROFILE_STEP=1 timeout 900 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/${R}_launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu --legs headline > $O/${R}_bench_ncu.log 2>&1
  # one frame of kazen's WarmStudio.xml at 512x512x64 (second frame of prof_paths.py)
  python tools/prof_paths.py warm > $O/${R}_prof_warm_plain.log 2>&1 &&
  timeout 300 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/${R}_launches_warm.csv python tools/prof_paths.py warm > /dev/null 2>&1
  # DRAM traffic of the traversal launches on the headline's own accel (10^8 triangles): first chunk of one 1920x1080x16 frame
  python tools/prof_paths.py big 100000000 1920 1080 16 > $O/${R}_prof_big100_plain.log 2>&1 &&
  KZ_PROF_LANES=1 timeout 600 $NCU --set full --profile-from-start off -k regex:k_extend -c 6 -o $O/prof_tmp python tools/prof_paths.py big 100000000 1920 1080 16 > /dev/null 2>&1
  python tools/ncu_summary.py $O/prof_tmp.ncu-rep > $O/${R}_big100_k_extend_summary.txt 2>&1; rm -f $O/prof_tmp.ncu-rep
  ;;
paths)
  # the wavefront kernels, full sections, single lane (serial kernels): WarmStudio 512x512x64 and the 10^7-triangle scene 1920x1080x16.
  # The first chunk of the profiled frame; reports are condensed here and deleted (no source import: the summaries are what is kept).
  python tools/prof_paths.py warm > $O/${R}_prof_warm_plain.log 2>&1 &&
  KZ_PROF_LANES=1 timeout 600 $NCU --set full --profile-from-start off -c 36 -o $O/prof_tmp python tools/prof_paths.py warm > $O/${R}_prof_warm_ncu.log 2>&1
  python tools/ncu_summary.py $O/prof_tmp.ncu-rep > $O/${R}_warm_kernels_summary.txt 2>&1; rm -f $O/prof_tmp.ncu-rep
  python tools/prof_paths.py big 10000000 1920 1080 16 > $O/${R}_prof_big_plain.log 2>&1 &&
  KZ_PROF_LANES=1 timeout 600 $NCU --set full --profile-from-start off -c 27 -o $O/prof_tmp python tools/prof_paths.py big 10000000 1920 1080 16 > $O/${R}_prof_big_ncu.log 2>&1
  python tools/ncu_summary.py $O/prof_tmp.ncu-rep > $O/${R}_big_kernels_summary.txt 2>&1; rm -f $O/prof_tmp.ncu-rep
  ;;
trace)
  # k_trace on the 2^20 (SAH) and 10^7 (LBVH) soups at the bench's batch size: launches 4 (primary) and 11 (incoherent) of variant_bench.py.
  # Each report (~5 MB with sources) is condensed here and deleted.
  M=_Z7k_trace7KzScenePK4KzF4jPfPjP9KzControl
  for cfg in "1m:" "10m:TRIS=10000000 LBVH=1"; do
    name=${cfg%%:*}; envs=${cfg#*:}
    env $envs RES=4096 NINC=16777216 python tools/variant_bench.py > $O/${R}_trace_${name}_plain.log 2>&1 || continue
    for which in "primary:4" "incoherent:11"; do
      w=${which%%:*}; skip=${which#*:}
      env $envs RES=4096 NINC=16777216 timeout 400 $NCU --set full --import-source on -k regex:k_trace -s $skip -c 1 -o $O/prof_tmp python tools/variant_bench.py > /dev/null 2>&1
      python tools/ncu_summary.py $O/prof_tmp.ncu-rep >> $O/${R}_k_trace_${name}_summary.txt 2>&1
      KZ_LAUNCH=0 python tools/ncu_lines.py $O/prof_tmp.ncu-rep k_trace $M 45 > $O/${R}_k_trace_${name}_lines_${w}.txt 2>&1
      rm -f $O/prof_tmp.ncu-rep
    done
  done
  ;;
*) echo "usage: $0 launches|paths|trace"; exit 2;;
esac
du -sh $O
