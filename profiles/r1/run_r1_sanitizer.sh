# compute-sanitizer on a subset of the GPU tests that reaches every kernel type (memcheck),
# plus racecheck on the shared-memory FFT kernels and the Bluestein path
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden and (nino3_morlet_tutorial or chirp4000_paul or chirp32k) or plan_classes or fp32_engine or resident_transform_products or test_unpadded_mode" 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_xwt_wct.py -x -q -k "wct_golden or smooth or seeded_exact or batch_channels or sharded" 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "chirp4000_morlet or plan_classes or test_unpadded_mode" 2>&1 | tail -4 ) | tee gpurun_out/sanitizer_r1.txt
