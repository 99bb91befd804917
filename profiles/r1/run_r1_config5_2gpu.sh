python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/config5_multi_gpu.py 512 2>&1 | grep -v Warning | tail -3
python profiles/config5_multi_gpu.py 256 2>&1 | grep -v Warning | tail -2
