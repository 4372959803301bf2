#!/bin/bash
# Final evidence of the round in one gpurun call: smoke, the bench lines of C1 / C2 / C3 / C5 (C4 at N = 1 costs 49 s per step: its line is
# taken separately), the reference arm, the C3 launch list of the same code, the config table.
set -x
o=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $o/r02_smoke.log 2>&1; tail -2 $o/r02_smoke.log
python bench.py --steps 3 --warmup 3 > $o/r02_bench_c3.json 2> $o/r02_bench_c3.err
python bench.py --workload c1 --steps 10 --warmup 3 > $o/r02_bench_c1.json 2> $o/r02_bench_c1.err
python bench.py --workload c2 --steps 5 --warmup 3 > $o/r02_bench_c2.json 2> $o/r02_bench_c2.err
python bench.py --workload c5 --steps 1 --warmup 3 --e2e-steps 1 > $o/r02_bench_c5.json 2> $o/r02_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/r02_bench_c3_reference_arm.json 2> $o/r02_bench_c3_reference_arm.err
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e > $o/r02_bench_c3_spp64.json 2> $o/r02_bench_c3_spp64.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $o/r02_launches_c3.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e > $o/r02_ncu_launches_c3.log 2>&1
python tools/config_table.py > $o/r02_config_table.md 2> $o/r02_config_table.err
cat $o/r02_config_table.md
