#!/bin/bash
# tools/build_variant.sh <out.so> [-DKNOB=VALUE ...]: one build of libkzgpu.so for tools/variant_bench.py; prints spills and the size of k_trace
out=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" -o $out /root/repo/nano-kazen_b200/csrc/kz_api.cu 2>&1 | grep -A2 "k_trace\|k_extend\|k_shadow\|k_occluded\|error" | grep "error\|spill" | head -30
cuobjdump -sass -fun '_Z7k_trace7KzScenePK4KzF4jPfPjP9KzControl' $out 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\s+\/\*\s*0x[0-9a-f]+\s*\*\/\s*$//' > ${out%.so}.k_trace.sass; wc -l ${out%.so}.k_trace.sass
