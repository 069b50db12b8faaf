#!/bin/bash
# GPU check of a round state: the GPU test suite, the default bench line, the drop-in's move by phase (outputs under gpurun_out/).
# usage: tools/round_check.sh [tag]
set -u; O=gpurun_out; T=${1:-r02u}
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest_rc=$?"; tail -4 $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench_rc=$?"; python tools/bench_summary.py $O/${T}_bench.json
python - <<PY
import json
d = json.load(open('$O/${T}_bench.json'))
print('dropin_config1', d['dropin_config1']['value'], '| fp8', d['fp8_tower']['value'], '| continuous', d['continuous_selfplay']['value'])
PY
python tools/dropin_profile.py 36 > $O/${T}_dropin_profile.txt 2>&1; cat $O/${T}_dropin_profile.txt
python tools/dropin_profile.py 200 > $O/${T}_dropin_profile200.txt 2>&1; cat $O/${T}_dropin_profile200.txt
