#!/bin/bash
# the two kernels added late in round 2 (cap_rows_kernel, tokenize_lookahead_kernel): full sections, one page
set -u; O=gpurun_out; T=${1:-r02}
ncu --set full --clock-control none -k "regex:cap_rows_kernel" -c 2 -f -o $O/${T}_extra1 python tools/all_kernels.py > $O/${T}_ncu_extra.log 2>&1
ncu --set full --clock-control none -k "regex:tokenize_lookahead_kernel" -c 2 -f -o $O/${T}_extra2 python tools/all_kernels.py 4096 lookahead >> $O/${T}_ncu_extra.log 2>&1
ncu -i $O/${T}_extra1.ncu-rep --page raw --csv > $O/${T}_extra_raw.csv 2> /dev/null
ncu -i $O/${T}_extra2.ncu-rep --page raw --csv 2> /dev/null | tail -n +3 >> $O/${T}_extra_raw.csv
rm -f $O/${T}_extra1.ncu-rep $O/${T}_extra2.ncu-rep
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference_rc=$?"; cat $O/${T}_bench_reference.json | cut -c1-600
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke_rc=$?"; tail -1 $O/${T}_smoke.log
wc -l $O/${T}_extra_raw.csv
