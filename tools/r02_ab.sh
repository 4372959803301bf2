#!/bin/bash
# A/B of two libptgpu builds inside one gpurun call: tools/r02_ab.sh <tag> <variant .so> [spp]
# Runs tools/profile_run.py (full C3, 8-spp passes) with the default build and with the variant, alternating twice.
tag=$1; var=$2; spp=${3:-8}
for i in 1 2; do
  echo "== default ($i)";  python tools/profile_run.py 200 100 1920 1080 $spp 2>&1 | tail -4
  echo "== $tag ($i)"; PTGPU_LIB=$var python tools/profile_run.py 200 100 1920 1080 $spp 2>&1 | tail -4
done
