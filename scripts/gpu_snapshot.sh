#!/bin/bash
# One gpurun call that refreshes the evidence for the current build (1 GPU):
#   scripts/gpu_snapshot.sh <tag> [full]
# writes gpurun_out/<tag>_*: GPU tests, the driver's bench command, secondary configs, the ncu launch list of a short
# bench run and (with "full") one --set full capture of the dominant kernels.  Numbers are never taken under ncu.
tag=${1:-snap}
mode=${2:-}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gputests.log
tail -3 $out/${tag}_gputests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench_configs.py 1 4 3 22 32 > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err; echo "configs rc=$?"
python bench_vecchia.py > $out/${tag}_vecchia.json 2> $out/${tag}_vecchia.err; echo "vecchia rc=$?"
python scripts/cold_start.py > $out/${tag}_cold.log 2>&1
short="python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline"
$short > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $short > $out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
if [ "$mode" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k 'regex:k_interp_cells|k_k8_finish|k_k8_scatter|k_gather|k_k8_stats' \
      -s 12 -c 6 -f -o $out/${tag}_prof $short > $out/${tag}_ncu2.log 2>&1
  echo "full capture rc=$?"
fi
