mkdir -p gpurun_out
{
for g in 0 2 4 64; do echo "#### CWTB_GEN_BAND=$g"; CWTB_GEN_BAND=$g timeout 300 python profiles/micro/config_kernels.py 3,5 2>&1; done
echo "#### profile"; timeout 300 python profiles/micro/config_kernels.py 3,5 --prof 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_cwt.py -x -q -m gpu -k "fp32 or config3 or config5 or f32 or batch" -s 2>&1 | grep -i "config\|passed\|failed\|error" | tail
} | tee gpurun_out/sweep_i.txt
