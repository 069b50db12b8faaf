#!/bin/bash
# Builds variants of libmcaz.so for an A/B on one GPU box (tools/ab/tower_ab.py): the built files travel with the gpurun
# snapshot.  usage: tools/ab/build_variants.sh name=<git-rev-or-.>:<extra nvcc flags> ...
set -eu
cd "$(dirname "$0")/../.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -shared -Xcompiler -fPIC -I include"
for spec in "$@"; do
    name=${spec%%=*}; rest=${spec#*=}; rev=${rest%%:*}; extra=${rest#*:}
    [ "$extra" = "$rest" ] && extra=""
    src=minitchess_alphazero_b200/csrc
    if [ "$rev" != "." ]; then
        tmp=$(mktemp -d); mkdir -p $tmp/csrc
        for f in $(git ls-tree --name-only $rev $src/); do git show $rev:$f > $tmp/csrc/$(basename $f); done
        git show $rev:include/mcaz.h > $tmp/mcaz.h
        nvcc ${FLAGS/-I include/-I $tmp} -I $tmp/csrc $extra $tmp/csrc/*.cu -o tools/ab/libmcaz_$name.so
        rm -rf $tmp
    else
        nvcc $FLAGS -I $src $extra $src/*.cu -o tools/ab/libmcaz_$name.so
    fi
    echo built tools/ab/libmcaz_$name.so
done
