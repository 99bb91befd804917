# 2 GPUs: bench under torchrun (config-2 replicas, config-5 channel shards with the NCCL gather
# through the C ABI, config-2 scale shards), reference arm under torchrun, sharded Monte-Carlo.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"   # + a fresh port per launch
timeout 900 $T 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; tail -c 1500 gpurun_out/bench_r2_2gpu.err | grep -v "^$" | tail -8
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_2gpu.json') if l.startswith('{')][-1])
print("N=2 value %.3e ms %.3f e2e %.3e (%.1f ms)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('topology'))
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, "value %.3e" % v['value'], "ms", v.get('ms_per_step'), "e2e", v.get('e2e'))
    else: print(k, v)
PY
timeout 300 $T 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 | cut -c1-300
timeout 600 $T 29513 profiles/config4_mc_multi_gpu.py 2>&1 | grep -v Warn | tail -6
timeout 600 $T 29514 profiles/config4_wct_sharded.py 2>&1 | grep -v Warn | tail -2
