#!/bin/bash
# Source-level ncu capture of one launch of a kernel (GPU box): per CUDA source line and per SASS instruction samples.
#   tools/capture_source.sh <tag> <kernel regex> <launch skip>
TAG=$1; RE=$2; SKIP=$3
OUT=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$RE" -s $SKIP -c 1 \
    -o /tmp/src_$TAG python tools/profile_step.py 36 1 > /dev/null 2>&1
ncu -i /tmp/src_$TAG.ncu-rep --page source --print-source cuda --csv > $OUT/src_${TAG}_cuda.csv 2>/dev/null
ncu -i /tmp/src_$TAG.ncu-rep --page source --print-source sass --csv > $OUT/src_${TAG}_sass.csv 2>/dev/null
ls -la $OUT/src_${TAG}_*.csv
