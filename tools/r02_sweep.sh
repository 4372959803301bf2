#!/bin/bash
# bench.py A/B sweep inside one gpurun call: each line "label|env assignments"; prints Msamples/s and Gpaths-bounce/s
spp=${SPP:-128}
run() { label=$1; shift; out=$(env "$@" python bench.py --spp $spp --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1); python - "$label" <<PY
import json,sys
d=json.loads('''$out''')
s=d["stage_ms_profiled_pass"]
print("%-28s %7.1f Msamples/s  %.4f Gpb/s  stages trace %.0f shade %.0f shadow %.0f mesh %.0f" % (sys.argv[1], d["value"], d["gpaths_bounce_per_s"], s["traceMs"], s["shadeMs"], s["shadowMs"], s["meshMs"]))
PY
}
V=ptsharp_b200/_lib/variants
run default A=1
run shade_order_off PTGPU_SHADE_ORDER=0
run queue_2^28 PTGPU_QUEUE_LOG2=28
run queue_2^26 PTGPU_QUEUE_LOG2=26
run node_burst_8 PTGPU_LIB=$V/libptgpu_nb8.so
run node_burst_2 PTGPU_LIB=$V/libptgpu_nb2.so
run leaf_burst_4 PTGPU_LIB=$V/libptgpu_lb4.so
run shade_sub_8 PTGPU_LIB=$V/libptgpu_sub8.so
run lanes_2 PTGPU_LANES=2
run default_again A=1
