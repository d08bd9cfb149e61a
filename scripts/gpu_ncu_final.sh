#!/bin/bash
tag=$1; out=gpurun_out; mkdir -p $out
short="python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline"
$short > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $short > $out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_interp_cells|k_k8_finish|k_k8_scatter|k_gather|k_k8_stats' \
    -s 12 -c 6 -f -o $out/${tag}_prof $short > $out/${tag}_ncu2.log 2>&1
echo "full capture rc=$?"
