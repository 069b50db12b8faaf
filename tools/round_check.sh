#!/bin/bash
# GPU check of a round state: the GPU test suite, the default bench line, the deferred-rows probe (outputs under gpurun_out/).  usage: tools/round_check.sh [tag]
set -u; O=gpurun_out; T=${1:-r02o}
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest_rc=$?"; tail -4 $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench_rc=$?"; python tools/bench_summary.py $O/${T}_bench.json
python tools/defer_probe.py 0 224 240 > $O/${T}_defer_probe.txt 2>&1; grep continuous $O/${T}_defer_probe.txt
