#!/bin/bash
# bf16 tower with one / two epilogue warp sets: tools/ab/build_variants.sh "base=.:" "new=.:-DTOWER_BF16_SETS=2" beforehand
set -u; O=gpurun_out; T=${1:-r02s2}
python tools/ab/tower_ab.py tools/ab/libmcaz_base.so tools/ab/libmcaz_new.so --rounds 2 > $O/${T}_sets_ab.txt 2>&1
B="python bench.py --no-e2e --no-cpu-baseline --no-plain --no-fp8 --steps 4 --warmup 3"
cp minitchess_alphazero_b200/libmcaz.so $O/libmcaz_shipped.so
for r in 1 2; do for v in base new; do
    cp tools/ab/libmcaz_$v.so minitchess_alphazero_b200/libmcaz.so
    $B > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
    python tools/bench_summary.py $O/${T}_bench_$v.json | sed "s/^/$v /" >> $O/${T}_sets_ab.txt
done; done
cp $O/libmcaz_shipped.so minitchess_alphazero_b200/libmcaz.so; rm -f $O/libmcaz_shipped.so
cut -c1-330 $O/${T}_sets_ab.txt
