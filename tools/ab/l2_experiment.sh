#!/bin/bash
# L2 residency experiment (round 2): evict_last policy on the tower's activation stores + persisting set-aside (TOWER_L2_MODE).
# Variants built beforehand: tools/ab/build_variants.sh "base=.:-DTOWER_L2_MODE=0" "new=.:" "newx=.:-DMCAZ_TIMING_EXPERIMENTS"
set -u
O=gpurun_out; T=${1:-r02k}
python tools/ab/tower_ab.py tools/ab/libmcaz_base.so tools/ab/libmcaz_new.so --rounds 2 > $O/${T}_ab.txt 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for v in base new; do
    # launch 9 = a warmed 2816-row forward, launch 420 = a 4096-row one (10 warm-ups + 400 timed forwards per size)
    ncu --metrics $M --clock-control none -k regex:tower_tc_kernel -s 8 -c 1 --csv --log-file $O/${T}_ncu_${v}_2816.csv \
        python tools/ab/tower_ab.py --worker tools/ab/libmcaz_$v.so > $O/${T}_ncu_$v.log 2>&1
    ncu --metrics $M --clock-control none -k regex:tower_tc_kernel -s 419 -c 1 --csv --log-file $O/${T}_ncu_${v}_4096.csv \
        python tools/ab/tower_ab.py --worker tools/ab/libmcaz_$v.so >> $O/${T}_ncu_$v.log 2>&1
done
for bp in 70 100; do
    echo "== MCAZ_L2_BUDGET_PCT=$bp" >> $O/${T}_ab.txt
    MCAZ_L2_BUDGET_PCT=$bp python tools/ab/tower_ab.py tools/ab/libmcaz_newx.so --rounds 1 >> $O/${T}_ab.txt 2>&1
    MCAZ_L2_BUDGET_PCT=$bp ncu --metrics $M --clock-control none -k regex:tower_tc_kernel -s 8 -c 1 --csv --log-file $O/${T}_ncu_newx${bp}_2816.csv \
        python tools/ab/tower_ab.py --worker tools/ab/libmcaz_newx.so > /dev/null 2>&1
done
# the whole step under the power cap: alternating
B="python bench.py --no-e2e --no-cpu-baseline --no-plain --no-fp8 --steps 4 --warmup 3"
cp minitchess_alphazero_b200/libmcaz.so $O/libmcaz_shipped.so
for r in 1 2 3; do for v in base new; do
    cp tools/ab/libmcaz_$v.so minitchess_alphazero_b200/libmcaz.so
    $B > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
    python tools/bench_summary.py $O/${T}_bench_$v.json | sed "s/^/$v /" >> $O/${T}_bench_ab.txt
done; done
cp $O/libmcaz_shipped.so minitchess_alphazero_b200/libmcaz.so; rm -f $O/libmcaz_shipped.so
python -m pytest tests/test_gpu_network.py tests/test_gpu_fp8.py -x -q -m gpu > $O/${T}_pytest.log 2>&1
cat $O/${T}_ab.txt; cat $O/${T}_bench_ab.txt; tail -3 $O/${T}_pytest.log
grep -h -A1 "dram__bytes\|gpu__time" /dev/null; for f in $O/${T}_ncu_*.csv; do echo $f; grep -E "dram__bytes|gpu__time|tensor" $f | awk -F'","' '{print "   " $(NF-2), $NF}'; done
