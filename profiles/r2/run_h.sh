mkdir -p gpurun_out
{
for g in 0 2 4 8 64; do echo "#### CWTB_GEN_BAND=$g"; CWTB_GEN_BAND=$g timeout 300 python profiles/micro/config_kernels.py 2,3,5 2>&1; done
echo "#### CWTB_EXPAND_MIN_R=2"; CWTB_EXPAND_MIN_R=2 timeout 300 python profiles/micro/config_kernels.py 3,5 2>&1
echo "#### profile"; timeout 300 python profiles/micro/config_kernels.py 3,4,5 --prof 2>&1
} | tee gpurun_out/sweep_h.txt
