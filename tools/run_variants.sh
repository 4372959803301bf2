#!/bin/bash
# usage: tools/run_variants.sh "<profile_run args>" name1 name2 ...   (runs on the GPU box)
args="$1"; shift
for v in "$@"; do
  echo "variant $v"
  PTGPU_LIB=ptsharp_b200/_lib/variants/libptgpu_$v.so timeout 200 python tools/profile_run.py $args 2>&1 | tail -2
done
