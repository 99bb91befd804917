mkdir -p gpurun_out
./profiles/micro/dmma_vs_dfma 2>&1 | tee gpurun_out/dmma_vs_dfma.txt
timeout 600 python profiles/micro/eps_tradeoff.py 2>&1 | tee gpurun_out/eps_tradeoff.txt
