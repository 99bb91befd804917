mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T profiles/config4_wct_sharded.py 2>&1 | grep -v Warn | tail -4 | tee gpurun_out/wct_sharded_2gpu.txt
