#!/bin/bash
spp=${SPP:-512}
run() { label=$1; shift; out=$(env "$@" python bench.py --spp $spp --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1); python - "$label" <<PY
import json,sys
d=json.loads('''$out''')
print("%-28s %7.1f Msamples/s  %.4f Gpb/s  %.1f ms/step" % (sys.argv[1], d["value"], d["gpaths_bounce_per_s"], d["ms_per_step"]))
PY
}
run default A=1
run lanes_2 PTGPU_LANES=2
run lanes_2_q28 PTGPU_LANES=2 PTGPU_QUEUE_LOG2=28
run lanes_3_q28 PTGPU_LANES=3 PTGPU_QUEUE_LOG2=28
run lanes_4_q28 PTGPU_LANES=4 PTGPU_QUEUE_LOG2=28
run lanes_4 PTGPU_LANES=4
