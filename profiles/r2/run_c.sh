mkdir -p gpurun_out
for p in 0 1 2; do echo "#### CWTB_PRIO=$p"; CWTB_PRIO=$p timeout 300 python profiles/micro/split_timing.py 2>&1; done | tee gpurun_out/split_c.txt
