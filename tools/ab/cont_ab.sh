#!/bin/bash
# continuous self-play A/B of two library builds (tools/ab/build_variants.sh "base=.:" "new=.:<flags>" beforehand), alternating
set -u; O=gpurun_out; T=${1:-r03c}
B="python bench.py --no-e2e --no-cpu-baseline --no-plain --no-fp8 --steps 4 --warmup 3 --mode continuous --free-sims 4"
cp minitchess_alphazero_b200/libmcaz.so $O/libmcaz_shipped.so
for r in 1 2; do for v in base new; do
    cp tools/ab/libmcaz_$v.so minitchess_alphazero_b200/libmcaz.so
    $B > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
    python tools/bench_summary.py $O/${T}_bench_$v.json | sed "s/^/$v /" >> $O/${T}_cont_ab.txt
done; done
cp $O/libmcaz_shipped.so minitchess_alphazero_b200/libmcaz.so; rm -f $O/libmcaz_shipped.so
cut -c1-250 $O/${T}_cont_ab.txt
