mkdir -p gpurun_out
{
echo "#### scalar (CWTB_EXPAND_MMA=0)"; CWTB_EXPAND_MMA=0 SPLIT_PARTS=full,expand timeout 300 python profiles/micro/split_timing.py 2>&1 | grep -v "=="
SPLIT_PARTS=full,expand timeout 600 python profiles/micro/split_timing.py pycwt_b200/variants/libcwtb200_l32.so pycwt_b200/variants/libcwtb200_l64.so pycwt_b200/libcwtb200.so pycwt_b200/variants/libcwtb200_l256.so 2>&1
timeout 300 python profiles/micro/config_kernels.py 2,4 --prof 2>&1 | grep "config\|Expand"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_cwt.py -x -q -m gpu -s 2>&1 | grep -i "config\|passed\|failed\|error" | tail -12
} | tee gpurun_out/sweep_q.txt
