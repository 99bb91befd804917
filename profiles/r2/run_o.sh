mkdir -p gpurun_out
{
for m in 0 1; do echo "#### CWTB_EXPAND_MMA=$m"; CWTB_EXPAND_MMA=$m SPLIT_PARTS=full,expand timeout 300 python profiles/micro/split_timing.py 2>&1 | grep -v "=="; CWTB_EXPAND_MMA=$m timeout 300 python profiles/micro/config_kernels.py 2,4 --prof 2>&1 | grep "config\|Expand"; done
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q -m gpu -s 2>&1 | grep -i "config\|passed\|failed\|error" | tail -12
} | tee gpurun_out/sweep_o.txt
