#!/bin/bash
# e4m3 tower A/B of the register split: tools/ab/build_variants.sh "base=HEAD:" "a=.:-DTOWER_FP8_TMEM_PIPE=0" "b=.:-DTOWER_FP8_TMEM_PIPE=1" beforehand
set -u; O=gpurun_out; T=${1:-r04a}
for f in 12 18; do
    echo "== e4m3 convolutions: $f" >> $O/${T}_fp8_ab.txt
    AB_FP8=$f python tools/ab/tower_ab.py tools/ab/libmcaz_base.so tools/ab/libmcaz_a.so tools/ab/libmcaz_b.so --rounds 2 >> $O/${T}_fp8_ab.txt 2>&1
done
cut -c1-330 $O/${T}_fp8_ab.txt
