mkdir -p gpurun_out
{
for a in 0 148; do for b in 0 148; do echo "#### CWTB_PF_ROWS_A=$a CWTB_PF_ROWS_B=$b"; CWTB_PF_ROWS_A=$a CWTB_PF_ROWS_B=$b timeout 300 python profiles/micro/config_kernels.py 4,5 2>&1; done; done
echo "#### CWTB_PF_ROWS_A=32 B=32"; CWTB_PF_ROWS_A=32 CWTB_PF_ROWS_B=32 timeout 300 python profiles/micro/config_kernels.py 4 2>&1
timeout 300 python profiles/micro/config_kernels.py 2,3 2>&1
} | tee gpurun_out/sweep_u.txt
