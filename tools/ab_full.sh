#!/bin/bash
# tools/ab_full.sh "c2 c3" lib1.so ...   like ab.sh, but keeps the per-ray node / triangle / instance counts
WL=$1; shift
for lib in "$@"; do for w in $WL; do RGB200_LIB=$PWD/$lib timeout -s KILL 90 python tools/gpu_time.py $w 2>&1 | sed "s#$PWD/##" | sed 's/ | primary.*//'; done; done
