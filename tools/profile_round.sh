#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun; outputs under gpurun_out/).  Every ncu pass follows a
# plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01c}
B="python bench.py --no-e2e --no-cpu-baseline --no-plain"
$B --sims 16 --steps 2 --warmup 3 > gpurun_out/${TAG}_plain16.json 2> gpurun_out/${TAG}_plain16.err || exit 1
# launch list (cold-cache, serialised: compare shares, not absolute times); skip the pre-roll and the warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --sims 16 --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt
# full sections of the tree / stem / head kernels inside a 200-simulation search
$B --sims 200 --steps 1 --warmup 1 > gpurun_out/${TAG}_plain200.json 2> gpurun_out/${TAG}_plain200.err || exit 1
ncu --set full --clock-control none --import-source on -k "regex:search_step_kernel|heads_legal_kernel|stem_onehot_kernel" -s 2400 -c 6 \
    -f -o gpurun_out/${TAG}_aux $B --sims 200 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_aux.log 2>&1
# the tower on a full 4096-row batch
python tools/net_bench.py 4096 10 > gpurun_out/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 5 -c 2 -f -o gpurun_out/${TAG}_tower \
    python tools/net_bench.py 4096 10 > gpurun_out/${TAG}_ncu_tower.log 2>&1
# the stateless rules kernels (BASELINE.json configs[1]): 1 M positions plain, then full sections on 65 536 positions
python tools/rules_bench.py > gpurun_out/${TAG}_rules_bench.json 2> gpurun_out/${TAG}_rules_bench.err || exit 1
ncu --set full --clock-control none --import-source on -k "regex:legal_moves_kernel|apply_kernel" -s 90 -c 2 -f -o gpurun_out/${TAG}_rules \
    python tools/rules_bench.py 65536 > gpurun_out/${TAG}_ncu_rules.log 2>&1
tail -n 2 gpurun_out/${TAG}_launches_summary.txt; tail -n 2 gpurun_out/${TAG}_ncu_aux.log; tail -n 2 gpurun_out/${TAG}_ncu_tower.log
