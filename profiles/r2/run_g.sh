mkdir -p gpurun_out
timeout 600 python profiles/micro/e2e_config4.py 2>&1 | grep -v "^$" | tee gpurun_out/e2e_config4.txt
