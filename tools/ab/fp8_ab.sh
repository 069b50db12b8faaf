#!/bin/bash
# e4m3 tower A/B on one box: tools/ab/build_variants.sh "base=HEAD:" "new=.:" beforehand
set -u; O=gpurun_out; T=${1:-r02q}
for f in 0 12 18; do
    echo "== e4m3 convolutions: $f" >> $O/${T}_fp8_ab.txt
    AB_FP8=$f python tools/ab/tower_ab.py tools/ab/libmcaz_base.so tools/ab/libmcaz_new.so --rounds 2 >> $O/${T}_fp8_ab.txt 2>&1
done
python -m pytest tests/test_gpu_fp8.py tests/test_gpu_network.py -x -q -m gpu > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
cat $O/${T}_fp8_ab.txt | cut -c1-330
