mkdir -p gpurun_out
{
for m in 0 1; do echo "#### CWTB_EXPAND_MMA=$m"; CWTB_EXPAND_MMA=$m timeout 300 python profiles/micro/config_kernels.py 3,5 --prof 2>&1 | grep "config\|Expand"; done
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q -m gpu -s -k "fp32 or config3 or config5 or f32 or batch or precision" 2>&1 | grep -i "config\|passed\|failed\|error" | tail -12
} | tee gpurun_out/sweep_x.txt
