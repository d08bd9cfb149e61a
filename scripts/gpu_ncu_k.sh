#!/bin/bash
# scripts/gpu_ncu_k.sh <tag> <kernel regex> [skip] [count]: one --set full capture of the named kernels in a short bench run
tag=$1; rx=$2; skip=${3:-4}; cnt=${4:-2}; out=gpurun_out; mkdir -p $out
short="python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline"
$short > $out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $out/${tag}_prof $short > $out/${tag}_ncu.log 2>&1
echo "capture rc=$?"
