#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus N): scripts/gpu_multi.sh <tag> <N>
tag=${1:-multi}; N=${2:-2}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_${N}gpu.txt 2>&1; echo "check(peer) rc=$?"
SK_COMM_TRANSPORT=nccl timeout 300 $TR tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_nccl_${N}gpu.txt 2>&1; echo "check(nccl) rc=$?"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench(peer) rc=$?"
SK_COMM_TRANSPORT=nccl timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_nccl_${N}gpu.json 2> $out/${tag}_bench_nccl_${N}gpu.err; echo "bench(nccl) rc=$?"
timeout 300 python tests/multi_gpu_group_check.py > $out/${tag}_group_check_${N}gpu.txt 2>&1; echo "group check rc=$?"
timeout 300 python bench_group.py --max-gpus $N > $out/${tag}_bench_group_${N}gpu.json 2> $out/${tag}_bench_group.err; echo "group bench rc=$?"
grep -h "multi_gpu" $out/${tag}_multi_gpu_check_${N}gpu.txt $out/${tag}_multi_gpu_check_nccl_${N}gpu.txt $out/${tag}_group_check_${N}gpu.txt
python - <<P
import json
for f in ("${tag}_bench_${N}gpu.json", "${tag}_bench_nccl_${N}gpu.json"):
    try:
        d = json.loads(open("$out/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", "%.3e" % d["value"], "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "strong ms", round(d.get("strong", {}).get("ms_per_step", 0), 4))
    except Exception as e:
        print(f, "unreadable", e)
P
