# usage: bash tools/ab_var.sh name1 name2 ...   ("product" = the in-tree library)
for v in "$@"; do
  if [ "$v" = product ]; then unset LRM_B200_LIB; else export LRM_B200_LIB=tools/_variants/liblrm_$v.so; fi
  python tools/tier_check.py ${AB_POINTS:-1000000000} lattice ${AB_CELL:-3} ${AB_DIM:-512} ${AB_KERNEL:-0} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/ab_$v.json'))
    print('$v', 'fused', round(d['three_tier']['fused_gpoints_s'],1), 'dist', round(d['three_tier']['dist_gpoints_s'],1), 'reach', round(d['three_tier']['reach_gpoints_s'],1), 'auto', round(d['auto']['fused_gpoints_s'],1), 'two_tier', round(d['two_tier']['fused_gpoints_s'],1), d['flags_equal'], d['max_vec_diff_mm'], d['auto_equal'], d['reach_equal'], 'first', round(d['three_tier']['first_call_s'],3), d['bricks_used'])
except Exception as e:
    print('$v', 'FAILED', e); print(open('gpurun_out/ab_$v.err').read()[-1500:])
"
done
