set -u; O=gpurun_out; T=r02m
timeout 900 python -m pytest tests/test_gpu_continuous.py tests/test_gpu_production_pin.py tests/test_gpu_lookahead.py tests/test_gpu_engine_edges.py -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest_rc=$?"; tail -15 $O/${T}_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-plain --no-fp8 --steps 4 --warmup 3"
for r in 1 2; do for d in 0 128 176; do
  $B --defer-rows $d > $O/${T}_bench_d$d.json 2> $O/${T}_bench_d$d.err; python tools/bench_summary.py $O/${T}_bench_d$d.json | sed "s/^/defer=$d lockstep /" >> $O/${T}_ab.txt
  $B --defer-rows $d --mode continuous --free-sims 4 > $O/${T}_benchc_d$d.json 2> $O/${T}_benchc_d$d.err; python tools/bench_summary.py $O/${T}_benchc_d$d.json | sed "s/^/defer=$d continuous /" >> $O/${T}_ab.txt
done; done
cat $O/${T}_ab.txt
