# correctness of the product library first (bounded: a deadlock must not hang the box), then the A/B
timeout 300 python -m pytest tests/test_one_leg_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/stress_tier.py 3 4200000 33000000 2>&1 | tail -1
for v in "$@"; do timeout 300 bash tools/ab_var.sh $v; done
