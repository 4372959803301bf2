#!/bin/bash
# A/B of libptgpu builds on C3 (8-spp passes of the full scene), alternating: tools/r02_c3_ab.sh rounds name1 name2 ...
n=$1; shift
for i in $(seq $n); do for v in "$@"; do
  lib=ptsharp_b200/_lib/variants/libptgpu_$v.so; [ "$v" = default ] && lib=ptsharp_b200/_lib/libptgpu.so
  echo "== $v ($i)"; PTGPU_LIB=$lib timeout 300 python tools/profile_cfg.py c3 8 2>&1 | tail -2
done; done
