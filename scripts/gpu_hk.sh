#!/bin/bash
tag=$1; out=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "hankel or bessel or dim" > $out/${tag}_hk_tests.log 2>&1; echo "tests rc=$?"; tail -2 $out/${tag}_hk_tests.log
python bench_configs.py 22 32 > $out/${tag}_hk_configs.jsonl 2>&1; echo "configs rc=$?"
python - <<P
import json
for l in open("$out/${tag}_hk_configs.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d["config"], "ms", round(d["ms"], 3), "interp_ms", round(d.get("interp_ms", 0), 3), "source_ms", round(d.get("source_ms", 0), 3), "err", d.get("max_err_over_k0"))
P
