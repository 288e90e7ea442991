#!/bin/bash
# full-set captures of the tcgen05 GEMM engine's instantiations (template arguments are only visible in demangled names)
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
STEP="python tools/profile_step.py 36 1"
cap() {   # name regex skip count
    ncu --set full --clock-control none --import-source off --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o /tmp/${TAG}_$1 $STEP > /dev/null 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1.csv 2>/dev/null
}
cap qkv_attn EpiAttn 32 32
cap proj EpiWindow 32 32
cap mlp_unfused "EpiRows<\(bool\)[01], \(bool\)[01]>, \(int\)[0-9]" 42 42
cap deembed AIm2col 1 1
cap split EpiSplit 9 9
cap frontend "AStftFrames|AIstft" 2 2
ls -la $OUT | tail -8
