mkdir -p gpurun_out
{
SPLIT_PARTS=full,band,exact,expand timeout 600 python profiles/micro/split_timing.py 2>&1
timeout 300 python profiles/micro/config_kernels.py 2,4 --prof 2>&1 | grep "config\|Expand\|PassA\|PassB"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s 2>&1 | grep -i "config\|passed\|failed\|error" | tail -12
} | tee gpurun_out/sweep_s.txt
