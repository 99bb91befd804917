mkdir -p gpurun_out
{
for f in 4 6 8; do echo "#### CWTB_PRIO_FAN=$f"; CWTB_PRIO_FAN=$f timeout 300 python profiles/micro/config_kernels.py 2,3,4 2>&1; done
} | tee gpurun_out/sweep_fan2.txt
