#!/bin/bash
# A/B of libptgpu builds on chosen configs: tools/r02_cfg_ab.sh "c2 16;c1 16;c3 8" name1 name2 ...
IFS=';' read -ra CFGS <<< "$1"; shift
for v in "$@"; do
  lib=ptsharp_b200/_lib/variants/libptgpu_$v.so; [ "$v" = default ] && lib=ptsharp_b200/_lib/libptgpu.so
  for cfg in "${CFGS[@]}"; do
    echo "== $v $cfg"; PTGPU_LIB=$lib timeout 300 python tools/profile_cfg.py $cfg 2>&1 | tail -2
  done
done
