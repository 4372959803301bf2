#!/bin/bash
spp=${SPP:-512}
run() { label=$1; shift; out=$(env "$@" python bench.py --spp $spp --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1); python - "$label" <<PY
import json,sys
d=json.loads('''$out''')
s=d["stage_ms_profiled_pass"]
print("%-28s %7.1f Msamples/s  %.4f Gpb/s  %.1f ms/step  stages trace %.0f shade %.0f shadow %.0f mesh %.0f" % (sys.argv[1], d["value"], d["gpaths_bounce_per_s"], d["ms_per_step"], s["traceMs"], s["shadeMs"], s["shadowMs"], s["meshMs"]))
PY
}
V=ptsharp_b200/_lib/variants
run default A=1
for v in "$@"; do run $v PTGPU_LIB=$V/libptgpu_$v.so; done
