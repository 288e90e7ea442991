#!/bin/bash
# Round-2 profile collection (run on the GPU box under gpurun, one GPU): writes CSV summaries under gpurun_out/.
#   tools/collect_profiles.sh [tag]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
STEP="python tools/profile_step.py 36 1"
# metric names that mention the tensor pipe / tcgen05 on this ncu build
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|utc|tmem|umma" | awk '{print $1}' | sort -u > $OUT/${TAG}_tensor_metric_names.txt
# 1. launch list of two steps (the second is the measured one): time + DRAM bytes of EVERY launch
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_dram.csv $STEP > /dev/null 2>&1
# 2. full-set captures of the distinct kernels, reduced to the raw metric page (CSV)
cap() {   # name regex skip count
    ncu --set full --clock-control none --import-source off -k regex:$2 -s $3 -c $4 -o /tmp/${TAG}_$1 $STEP > /dev/null 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1.csv 2>/dev/null
}
cap mlp_fused mlp_fused 16 16
cap qkv_attn EpiAttn 32 32
cap proj EpiWindow 32 32
cap mlp_unfused "EpiRows" 40 40
cap deembed AIm2col 1 1
cap pvq_stream pvq_stream 12 12
cap conv3x3 conv3x3 1 1
cap patch_embed patch_embed 1 1
cap frontend "AStftFrames|AIstft" 2 2
ls -la $OUT | tail -20
