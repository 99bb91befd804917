mkdir -p gpurun_out
{
timeout 600 python profiles/micro/eps_tradeoff.py 2>&1
export SPLIT_PARTS=full,exact
for m in 0 1 2; do echo "#### CWTB_DENSE_MARGIN=$m"; CWTB_DENSE_MARGIN=$m timeout 300 python profiles/micro/split_timing.py 2>&1 | grep -v "=="; done
for ch in 2 4; do for g in 2 3 4 8 16; do echo "#### CWTB_CHAINS=$ch CWTB_GROUP=$g"; CWTB_CHAINS=$ch CWTB_GROUP=$g timeout 300 python profiles/micro/split_timing.py 2>&1 | grep -v "=="; done; done
} | tee gpurun_out/sweep_e.txt
