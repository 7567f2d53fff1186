#!/bin/bash
# The two bench lines at N GPUs (charged N x): gpurun --gpus 8 --timeout 400 -- 'bash tools/multi_gpu_short.sh 8'
N=${1:-8}
O=gpurun_out/multi_${N}
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_weak.json 2> $O/bench_weak.err; echo "weak rc=$?"
timeout 120 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-posit > $O/bench_strong.json 2> $O/bench_strong.err; echo "strong rc=$?"
python - <<PY
import json
for f in ("bench_weak", "bench_strong"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["per_rank"]["kernel_ms"], d["clocks"], d.get("positionability", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
