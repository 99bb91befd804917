mkdir -p gpurun_out
{
for g in 0 2 4; do echo "#### CWTB_GEN_BAND=$g"; CWTB_GEN_BAND=$g timeout 300 python profiles/micro/config_kernels.py 2,3,5 2>&1; done
echo "#### profile"; timeout 300 python profiles/micro/config_kernels.py 5 --prof 2>&1 | grep -v "coarse:\|fwd:"
} | tee gpurun_out/sweep_k.txt
