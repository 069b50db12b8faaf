#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun; outputs under gpurun_out/).  Every ncu pass follows a
# plain run of the same command that exited 0; numbers printed under ncu are never bench values.  The .ncu-rep files are
# turned into raw CSV pages on the box and deleted (gpurun brings back at most 64 MiB).
set -u
TAG=${1:-r02}
O=gpurun_out
page() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2> /dev/null; rm -f $O/$1.ncu-rep; }
B="python bench.py --no-e2e --no-cpu-baseline --no-plain --no-fp8"
$B --sims 16 --steps 2 --warmup 3 > $O/${TAG}_plain16.json 2> $O/${TAG}_plain16.err || exit 1
# launch list (cold-cache, serialised: compare shares, not absolute times); skip the pre-roll and the warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 300 --csv --log-file $O/${TAG}_launches.csv \
    $B --sims 16 --steps 2 --warmup 3 > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launches_summary.txt
# full sections of the tree / stem / head / tower kernels inside a 200-simulation search of the default workload
$B --sims 200 --steps 1 --warmup 1 > $O/${TAG}_plain200.json 2> $O/${TAG}_plain200.err || exit 1
ncu --set full --clock-control none --import-source on -k "regex:search_step_kernel|heads_legal_kernel|stem_onehot_kernel|tower_tc_kernel" -s 2400 -c 8 \
    -f -o $O/${TAG}_hot $B --sims 200 --steps 1 --warmup 1 > $O/${TAG}_ncu_hot.log 2>&1
page ${TAG}_hot
# the tower alone: a full 4096-row batch, the bench's typical 2816 rows, and the e4m3 form
python tools/net_bench.py 4096 10 > $O/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 5 -c 2 -f -o $O/${TAG}_tower4096 \
    python tools/net_bench.py 4096 10 > $O/${TAG}_ncu_tower.log 2>&1
page ${TAG}_tower4096
python tools/net_bench.py 2816 10 >> $O/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 5 -c 2 -f -o $O/${TAG}_tower2816 \
    python tools/net_bench.py 2816 10 >> $O/${TAG}_ncu_tower.log 2>&1
ncu -i $O/${TAG}_tower2816.ncu-rep --page raw --csv > $O/${TAG}_tower2816_raw.csv 2> /dev/null       # this one report is kept (9 MB: source page)
python tools/net_bench.py 2816 10 12 >> $O/${TAG}_netbench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 7 -c 2 -f -o $O/${TAG}_tower2816_fp8 \
    python tools/net_bench.py 2816 10 12 >> $O/${TAG}_ncu_tower.log 2>&1
page ${TAG}_tower2816_fp8
# every other kernel of the library at working sizes (two launches of each are enough: -c counts all matching launches)
python tools/all_kernels.py > $O/${TAG}_all_kernels.log 2>&1 || exit 1
ncu --set full --clock-control none \
    -k "regex:legal_moves_kernel|apply_kernel|tokenize_kernel|perft|play_device_kernel|restart_finished_kernel|recycle_kernel|root_stats_kernel|game_states_kernel|node_stats_kernel|reset_games_kernel|reset_trees_kernel|set_positions_kernel|play_kernel|heads_kernel|collate_kernel|sample_root_noise_kernel|select_expand_kernel|backup_kernel|untag_rows_kernel|cap_rows_kernel|prep_|calib_positions" \
    -c 90 -f -o $O/${TAG}_all python tools/all_kernels.py > $O/${TAG}_ncu_all.log 2>&1
page ${TAG}_all
ncu --set full --clock-control none -k "regex:search_step_kernel|heads_legal_kernel|tokenize_lookahead_kernel" -s 3 -c 9 -f -o $O/${TAG}_lookahead \
    python tools/all_kernels.py 4096 lookahead > $O/${TAG}_ncu_lookahead.log 2>&1
page ${TAG}_lookahead
# wait statistics of the tower (bf16 at 4096 / 2816 / 256 rows, e4m3 at 2816), the host-buffer step by phase, the drop-in's move by phase
for r in 4096 2816 256; do MCAZ_TOWER_STATS=1 python tools/tower_stats.py $r; done > $O/${TAG}_tower_wait_stats.txt 2>&1
MCAZ_TOWER_STATS=1 python tools/tower_stats.py 2816 fp8 >> $O/${TAG}_tower_wait_stats.txt 2>&1
python tools/e2e_phases.py > $O/${TAG}_e2e_phases.txt 2>&1
python tools/dropin_profile.py 36 > $O/${TAG}_dropin_profile.txt 2>&1
python tools/kernel_table.py $O/${TAG}_hot_raw.csv $O/${TAG}_tower4096_raw.csv $O/${TAG}_tower2816_raw.csv $O/${TAG}_tower2816_fp8_raw.csv $O/${TAG}_lookahead_raw.csv \
    $O/${TAG}_all_raw.csv > $O/${TAG}_kernel_table.md
ls -la $O | grep ${TAG}_
