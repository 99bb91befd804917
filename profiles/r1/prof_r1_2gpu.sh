mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1
  timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1
  timeout 600 $TR --master-port 29513 profiles/config5_multi_gpu.py 512 2>&1 | grep config5
  timeout 600 $TR --master-port 29514 profiles/config4_mc_multi_gpu.py 200 2>&1 | grep config4
  timeout 600 python profiles/config4_mc_multi_gpu.py 200 2>&1 | grep config4 ) | tee gpurun_out/multi_gpu_r1.txt | cut -c1-600
