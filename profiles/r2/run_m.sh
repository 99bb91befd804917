mkdir -p gpurun_out
{
SPLIT_PARTS=full,dense,band,exact timeout 600 python profiles/micro/split_timing.py pycwt_b200/libcwtb200.so pycwt_b200/variants/libcwtb200_b4.so pycwt_b200/variants/libcwtb200_b16.so pycwt_b200/variants/libcwtb200_a8.so 2>&1
timeout 300 python profiles/micro/config_kernels.py 3,4,5 2>&1
} | tee gpurun_out/sweep_m.txt
