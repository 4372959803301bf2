#!/bin/bash
# ncu --set full of k_scene_trace<START> launches of the measured C4 pass: tools/r02_ncu_scene.sh <skip> <count> variant...
skip=$1; count=$2; shift; shift
for v in "$@"; do
  lib=ptsharp_b200/_lib/variants/libptgpu_$v.so; [ "$v" = default ] && lib=ptsharp_b200/_lib/libptgpu.so
  PTGPU_LIB=$lib timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'k_scene_trace<\(int\)0>' -s $skip -c $count -f -o gpurun_out/prof_start_c4_$v python tools/c4_detail.py 2 > gpurun_out/ncu_start_$v.log 2>&1
  tail -3 gpurun_out/ncu_start_$v.log
done
