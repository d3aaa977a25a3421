#!/bin/bash
# tools/ab.sh "c2 c3" lib1.so lib2.so ...   (developer A/B timing on the GPU box)
WL=$1; shift
for lib in "$@"; do for w in $WL; do RGB200_LIB=$PWD/$lib timeout -s KILL 90 python tools/gpu_time.py $w 2>&1 | grep -v "per ray" | sed "s#$PWD/##"; done; done
