for b in 1 0; do
LRM_TC_BRICKS=$b python tools/run_fused.py 400000000 3 both 1 > gpurun_out/ncu_b$b.plain.log 2>&1 || exit 1
LRM_TC_BRICKS=$b ncu --set full --clock-control none --import-source on -k regex:one_leg_tier_kernel --launch-skip 1 -c 1 -f -o gpurun_out/tier_b$b python tools/run_fused.py 400000000 3 both 1 > gpurun_out/ncu_b$b.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
