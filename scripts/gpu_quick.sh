#!/bin/bash
# scripts/gpu_quick.sh <tag>: all GPU tests, the host timeline of a step, and the driver's bench command (1 GPU)
tag=${1:-q}; out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_gputests.log
python scripts/step_timeline.py > $out/${tag}_timeline.log 2>&1; echo "timeline rc=$?"; cat $out/${tag}_timeline.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<P
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 4), "interp", round(d["interp_ms_per_step"], 4), "sort", round(d["sort_ms_per_step"], 4),
      "gather", round(d["gather_ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), "step frac", round(d["roofline"]["step"]["frac"], 4),
      "K4 frac", round(d["roofline"]["frac"], 3), "clocks", d["clocks"])
P
