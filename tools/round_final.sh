#!/bin/bash
# last GPU pass of a round: GPU tests, smoke, the default bench line, the e4m3 tower's page and wait statistics (outputs under gpurun_out/)
set -u; O=gpurun_out; T=${1:-r02z}
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest_rc=$?"; tail -3 $O/${T}_pytest.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke_rc=$?"
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench_rc=$?"; python tools/bench_summary.py $O/${T}_bench.json
python - <<PY
import json
d = json.load(open('$O/${T}_bench.json'))
print('dropin_config1', d['dropin_config1']['value'], '| fp8', d['fp8_tower']['value'], d['fp8_tower']['roofline']['mma_frac'], d['fp8_tower']['all_18_convolutions']['value'], '| continuous', d['continuous_selfplay']['value'])
PY
python tools/net_bench.py 2816 10 12 > $O/${T}_netbench_fp8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tower_tc_kernel -s 7 -c 2 -f -o $O/${T}_tower2816_fp8 python tools/net_bench.py 2816 10 12 > $O/${T}_ncu_fp8.log 2>&1
ncu -i $O/${T}_tower2816_fp8.ncu-rep --page raw --csv > $O/${T}_tower2816_fp8_raw.csv 2> /dev/null
MCAZ_TOWER_STATS=1 python tools/tower_stats.py 2816 fp8 > $O/${T}_tower_wait_stats_fp8.txt 2>&1; cat $O/${T}_tower_wait_stats_fp8.txt
