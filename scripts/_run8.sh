STEP="python scripts/profile_step.py 64e6 both"
VR_LIB_PATH=$PWD/variants/hyb.so ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 2 -c 1 -f -o gpurun_out/prof_r2h_shade_hyb $STEP > gpurun_out/ncu_r2h.log 2>&1
VR_LIB_PATH=$PWD/variants/r1rng.so ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 2 -c 1 -f -o gpurun_out/prof_r2h_shade_r1rng $STEP > gpurun_out/ncu_r2h_b.log 2>&1
ls -la gpurun_out/prof_r2h*
