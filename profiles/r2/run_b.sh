mkdir -p gpurun_out
timeout 600 python profiles/micro/split_timing.py pycwt_b200/variants/libcwtb200_noasync.so pycwt_b200/libcwtb200.so 2>&1 | tee gpurun_out/split_b.txt
timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
