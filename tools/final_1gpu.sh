# one single-GPU visit: the bench lines, the launch list and the full capture of the fused kernel
O=gpurun_out/final1
mkdir -p $O
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench_weak.json 2> $O/bench_weak.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --scaling strong --no-posit --no-cpu > $O/bench_strong.json 2> $O/bench_strong.err; echo "strong rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-posit > $O/ncu_launch.log 2>&1; echo "launches rc=$?"
NCU_POINTS=1000000000 bash tools/ncu_one.sh final > $O/ncu_full.log 2>&1; echo "full rc=$?"
tail -c 300 $O/bench_weak.json
