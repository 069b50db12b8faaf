#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun; outputs under gpurun_out/).  Every ncu pass follows a
# plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02}
B="python bench.py --no-e2e --no-cpu-baseline --no-plain"
$B --sims 16 --steps 2 --warmup 3 > gpurun_out/${TAG}_plain16.json 2> gpurun_out/${TAG}_plain16.err || exit 1
# launch list (cold-cache, serialised: compare shares, not absolute times); skip the pre-roll and the warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --sims 16 --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt
# full sections of the tree / stem / head / tower kernels inside a 200-simulation search of the default workload
$B --sims 200 --steps 1 --warmup 1 > gpurun_out/${TAG}_plain200.json 2> gpurun_out/${TAG}_plain200.err || exit 1
ncu --set full --clock-control none --import-source on -k "regex:search_step_kernel|heads_legal_kernel|stem_onehot_kernel|tower_tc_kernel" -s 2400 -c 8 \
    -f -o gpurun_out/${TAG}_hot $B --sims 200 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_hot.log 2>&1
# the tower on a full 4096-row batch and on the bench's typical 2816 rows
python tools/net_bench.py 4096 10 > gpurun_out/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 5 -c 2 -f -o gpurun_out/${TAG}_tower4096 \
    python tools/net_bench.py 4096 10 > gpurun_out/${TAG}_ncu_tower.log 2>&1
python tools/net_bench.py 2816 10 >> gpurun_out/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 5 -c 2 -f -o gpurun_out/${TAG}_tower2816 \
    python tools/net_bench.py 2816 10 >> gpurun_out/${TAG}_ncu_tower.log 2>&1
# every other kernel of the library at working sizes
python tools/all_kernels.py > gpurun_out/${TAG}_all_kernels.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    -k "regex:legal_moves_kernel|apply_kernel|tokenize_kernel|perft|play_device_kernel|restart_finished_kernel|recycle_kernel|root_stats_kernel|game_states_kernel|node_stats_kernel|reset_games_kernel|reset_trees_kernel|set_positions_kernel|play_kernel|heads_kernel|collate_kernel|sample_root_noise_kernel|select_expand_kernel|backup_kernel|untag_rows_kernel|prep_" \
    -c 170 -f -o gpurun_out/${TAG}_all python tools/all_kernels.py > gpurun_out/${TAG}_ncu_all.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:search_step_kernel|heads_legal_kernel" -s 2 -c 6 -f -o gpurun_out/${TAG}_lookahead \
    python tools/all_kernels.py 4096 lookahead > gpurun_out/${TAG}_ncu_lookahead.log 2>&1
tail -n 2 gpurun_out/${TAG}_launches_summary.txt; tail -n 2 gpurun_out/${TAG}_ncu_hot.log; tail -n 2 gpurun_out/${TAG}_ncu_all.log
