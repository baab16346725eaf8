mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_matrix.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/bt_tests.log
for v in cnt bt; do
  echo "== $v" >> gpurun_out/bt_perf.log
  VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_step.py 256e6 both 2>&1 | tail -1 >> gpurun_out/bt_perf.log
  VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_c5.py 100e6 2>&1 | grep "rep 1" >> gpurun_out/bt_perf.log
done
