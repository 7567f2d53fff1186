set -x
for b in 1 0; do LRM_TC_BRICKS=$b python tools/tier_check.py 1000000000 lattice 3 512 0 > gpurun_out/bricks_$b.json 2> gpurun_out/bricks_$b.err; done
python -m pytest tests/test_one_leg_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/bricks_tests.log
cat gpurun_out/bricks_tests.log
for f in 1 0; do python -c "
import json,sys
d=json.load(open('gpurun_out/bricks_$f.json'))
print('bricks=$f', round(d['three_tier']['fused_gpoints_s'],1), round(d['three_tier']['dist_gpoints_s'],1), round(d['three_tier']['reach_gpoints_s'],1), round(d['auto']['fused_gpoints_s'],1), d['flags_equal'], d['max_vec_diff_mm'], d['auto_equal'], d['reach_equal'], d['three_tier']['first_call_s'], d['bricks_used'], d['brick_capacity'])
"; done
tail -3 gpurun_out/bricks_1.err
