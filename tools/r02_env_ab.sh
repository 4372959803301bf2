#!/bin/bash
# A/B of an environment switch of libptgpu: tools/r02_env_ab.sh VAR "v0 v1" rounds "c3 8;c4 2"
var=$1; vals=$2; n=$3; IFS=';' read -ra CFGS <<< "$4"
for i in $(seq $n); do for v in $vals; do for cfg in "${CFGS[@]}"; do
  echo "== $var=$v $cfg ($i)"; env $var=$v timeout 300 python tools/profile_cfg.py $cfg 2>&1 | tail -2
done; done; done
