#!/bin/bash
# Round-2 evidence in one gpurun call: the five BASELINE configs through bench.py, then (each after its command exited 0 without ncu)
# the ncu launch list of a C3 bench step and full captures of k_mesh / k_march.  Everything lands in gpurun_out/.
set -x
o=gpurun_out
python bench.py --steps 3 --warmup 3 > $o/r02_bench_c3.json 2> $o/r02_bench_c3.err
python bench.py --workload c1 --steps 10 --warmup 3 > $o/r02_bench_c1.json 2> $o/r02_bench_c1.err
python bench.py --workload c2 --steps 5 --warmup 3 > $o/r02_bench_c2.json 2> $o/r02_bench_c2.err
python bench.py --workload c5 --steps 1 --warmup 3 --e2e-steps 1 > $o/r02_bench_c5.json 2> $o/r02_bench_c5.err
python bench.py --workload c4 --steps 1 --warmup 3 --e2e-steps 1 > $o/r02_bench_c4.json 2> $o/r02_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/r02_bench_c3_reference_arm.json 2> $o/r02_bench_c3_reference_arm.err
# launch list of one 64-spp C3 step (the batch size of the 512-spp bench)
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e > $o/r02_bench_c3_spp64.json 2> $o/r02_bench_c3_spp64.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $o/r02_launches_c3.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e > $o/r02_ncu_launches_c3.log 2>&1
# full captures: k_mesh on C3 (first five launches of a 2-spp pass: depth-0 trace rounds 0/1, depth-0 shadow rounds 0/1, depth-1 trace round 0)
python tools/profile_run.py 200 100 1920 1080 2 > $o/r02_profile_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mesh -s 0 -c 5 -f -o $o/prof_mesh_r02 python tools/profile_run.py 200 100 1920 1080 2 > $o/r02_ncu_mesh.log 2>&1
# k_march on C5 (first two launches of each kind) and the C5 launch list
python tools/profile_cfg.py c5 2 > $o/r02_profile_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_march -s 0 -c 4 -f -o $o/prof_march_r02 python tools/profile_cfg.py c5 2 > $o/r02_ncu_march.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/r02_launches_c5.csv python tools/profile_cfg.py c5 2 > $o/r02_ncu_launches_c5.log 2>&1
tail -c 400 $o/r02_bench_c3.json
