#!/bin/bash
# usage: tools/cfg_variants.sh "<config_table args>" name1 name2 ...   (runs on the GPU box; variants from tools/build_variants.py)
args="$1"; shift
for v in "$@"; do
  echo "variant $v"
  PTGPU_LIB=ptsharp_b200/_lib/variants/libptgpu_$v.so timeout 300 python tools/config_table.py $args 2>&1 | tail -n +3
done
