#!/bin/bash
# A/B of libptgpu builds on the scene level (C4 = instanced, C3, C5): tools/r02_scene_ab.sh name1 name2 ...  ("default" = the product build)
for v in "$@"; do
  lib=ptsharp_b200/_lib/variants/libptgpu_$v.so; [ "$v" = default ] && lib=ptsharp_b200/_lib/libptgpu.so
  for cfg in "c4 2" "c3 8" "c5 2"; do
    echo "== $v $cfg"; PTGPU_LIB=$lib timeout 300 python tools/profile_cfg.py $cfg 2>&1 | tail -2
  done
done
