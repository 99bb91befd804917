# compute-sanitizer on a subset of the GPU tests that reaches every kernel type of the round-2 build
# (memcheck), plus racecheck on the shared-memory kernels: tile FFTs, tensor-core expansion (both
# column layouts), coherence kernels, pipelined batch
mkdir -p gpurun_out
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden and (nino3_morlet_tutorial or chirp4000_paul or chirp32k) or plan_classes or expansion_path or expansion_by_four or fp32_engine or resident_transform_products or test_unpadded_mode" 2>&1 | tail -4
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_xwt_wct.py -x -q -k "wct_golden or smooth_golden or seeded_exact or batch_channels or medium_two_kernel or sharded" 2>&1 | tail -4
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q -k "chirp4000_morlet or plan_classes or expansion_by_four or test_unpadded_mode or wct_golden or batch_channels" 2>&1 | tail -4 ) | tee gpurun_out/sanitizer_r2.txt
