#!/bin/bash
# One 8-GPU box visit (charged 8x): everything that needs N > 1, each leg under its own timeout.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/multi_gpu_round.sh 8'
N=${1:-8}
O=gpurun_out/multi_${N}
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_weak.json 2> $O/bench_weak.err; echo "weak rc=$?"
timeout 120 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-posit > $O/bench_strong.json 2> $O/bench_strong.err; echo "strong rc=$?"
timeout 120 $TR tools/bench_posit.py --config c5 --check 500 > $O/posit_c5.json 2> $O/posit_c5.err; echo "c5 rc=$?"
timeout 200 $TR tools/bench_posit.py --config c4 --check 100 > $O/posit_c4.json 2> $O/posit_c4.err; echo "c4 rc=$?"
[ -n "$SKIP_PCIE" ] || timeout 90 $TR tools/pcie_probe.py > $O/pcie.json 2> $O/pcie.err; echo "pcie rc=$?"
tail -c 600 $O/bench_weak.json; echo; tail -c 400 $O/bench_strong.json; echo; cat $O/posit_c5.json $O/posit_c4.json | cut -c1-700
