#!/bin/bash
# peer-transport only (bitwise check + bench), N GPUs: scripts/gpu_multi_peer.sh <tag> <N>
tag=${1:-mp}; N=${2:-2}; out=gpurun_out; mkdir -p $out
export SK_PEER_TIMEOUT_S=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 240 $TR tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_${N}gpu.txt 2>&1; echo "check(peer, chained) rc=$?"
SK_SHARDED_CHAIN=0 timeout 240 $TR tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_nochain_${N}gpu.txt 2>&1; echo "check(peer, no chain) rc=$?"
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench(chained) rc=$?"
SK_SHARDED_CHAIN=0 timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-strong-leg > $out/${tag}_bench_nochain_${N}gpu.json 2> $out/${tag}_bench_nochain_${N}gpu.err; echo "bench(no chain) rc=$?"
grep -h "multi_gpu\|Error\|error" $out/${tag}_multi_gpu_check_${N}gpu.txt $out/${tag}_multi_gpu_check_nochain_${N}gpu.txt | head -20
tail -3 $out/${tag}_bench_${N}gpu.err
python - <<P
import json
for f in ("${tag}_bench_${N}gpu.json", "${tag}_bench_nochain_${N}gpu.json"):
    try:
        d = json.loads(open("$out/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", "%.3e" % d["value"], "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "strong ms", round(d.get("strong", {}).get("ms_per_step", 0), 4), "err", d["parity"]["max_abs_err_vs_closed_form"])
    except Exception as e:
        print(f, "unreadable", e)
P
