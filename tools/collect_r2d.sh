#!/bin/bash
# Final evidence pass of round 2 (GPU box, one GPU): tests, bench, sanitizers, launch list with DRAM bytes, full-set ncu
# captures of the kernels that changed after tools/collect_r2b.sh (fused attention at C = 45, output conv, patch embed).
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -2 $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
STEP="python tools/profile_step.py 36 1"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_dram.csv $STEP > /dev/null 2>&1
cap() {   # name regex skip count
    timeout 600 ncu --set full --clock-control none --import-source off --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o /tmp/${TAG}_$1 $STEP > /dev/null 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1.csv 2>/dev/null
}
cap qkv_attn EpiAttn 32 32
cap conv3x3 conv3x3 1 1
cap patch_embed patch_embed 1 1
tail -n 2 $OUT/${TAG}_sanitizer_memcheck.log $OUT/${TAG}_sanitizer_racecheck.log
ls -la $OUT | grep ${TAG}_
