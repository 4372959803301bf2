#!/bin/bash
# A/B of libptgpu builds on C5 (2-spp passes): tools/r02_c5_ab.sh name1 name2 ...
for v in "$@"; do
  lib=ptsharp_b200/_lib/variants/libptgpu_$v.so; [ "$v" = default ] && lib=ptsharp_b200/_lib/libptgpu.so
  echo "== $v"; PTGPU_LIB=$lib timeout 300 python tools/profile_cfg.py c5 2 2>&1 | tail -2
done
