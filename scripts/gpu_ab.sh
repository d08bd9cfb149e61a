#!/bin/bash
# quick A/B on one GPU: scripts/gpu_ab.sh <tag> "<ENV=val ...>" ["<ENV=val ...>" ...]   (bench.py without the CPU leg)
tag=$1; shift; out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -x -q -k "sort_paths or full_size or interp_modes" > $out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $out/${tag}_tests.log
i=0
for envs in "$@"; do
  env $envs python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_ab$i.json 2> $out/${tag}_ab$i.err
  python - <<P
import json
d = json.loads(open("$out/${tag}_ab$i.json").read().strip().splitlines()[-1])
print("[$envs]", "ms/step", round(d["ms_per_step"], 4), "interp", round(d["interp_ms_per_step"], 4), "sort", round(d["sort_ms_per_step"], 4),
      "gather", round(d["gather_ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), "step frac", round(d["roofline"]["step"]["frac"], 4))
P
  i=$((i+1))
done
