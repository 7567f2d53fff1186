# usage: bash tools/ncu_one.sh name   -> gpurun_out/tier_<name>.ncu-rep of the product (or LRM_B200_LIB) library
ncu --set full --clock-control none --import-source on -k regex:one_leg_tier_kernel --launch-skip 1 -c 1 -f -o gpurun_out/tier_$1 python tools/run_fused.py ${NCU_POINTS:-400000000} 3 both 1 > gpurun_out/ncu_$1.log 2>&1
ls -la gpurun_out/tier_$1.ncu-rep
