#!/bin/bash
# Evidence of the final round-2 build (run on the GPU box under gpurun, one GPU): sanitizer logs, the launch list with
# DRAM bytes of one step, full-set ncu captures of every distinct kernel (raw metric page as CSV) -> gpurun_out/.
#   tools/collect_r2b.sh [tag]
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
STEP="python tools/profile_step.py 36 1"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_dram.csv $STEP > /dev/null 2>&1
cap() {   # name regex skip count
    timeout 600 ncu --set full --clock-control none --import-source off --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o /tmp/${TAG}_$1 $STEP > /dev/null 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1.csv 2>/dev/null
}
cap qkv_attn EpiAttn 32 32
cap mlp_fused mlp_fused 16 16
cap proj EpiWindow 32 32
cap mlp_unfused "EpiRows<\(bool\)[01], \(bool\)[01]>, \(int\)[0-9]" 30 30
cap deembed AIm2col 1 1
cap split EpiSplit 9 9
cap pvq_stream pvq_stream 12 12
cap conv3x3 conv3x3 1 1
cap patch_embed patch_embed 1 1
cap frontend "AStftFrames|AIstft" 2 2
tail -3 $OUT/${TAG}_sanitizer_memcheck.log $OUT/${TAG}_sanitizer_racecheck.log
ls -la $OUT | grep ${TAG}_
