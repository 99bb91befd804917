# 8 GPUs of one box: the bench line the driver's scaling run produces (config-2 replicas, config-5 channel
# shards with the NCCL gather through the C ABI, config-2 scale shards)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $T bench.py --gpus 8 --steps 10 --warmup 3 --configs 5 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; tail -c 1200 gpurun_out/bench_r2_8gpu.err | grep -v "^$" | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_8gpu.json') if l.startswith('{')][-1])
print("N=8 value %.3e ms %.3f e2e %.3e (%.1f ms per rank)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('topology'), d.get('clocks'))
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, "value %.3e" % v['value'], "ms", v.get('ms_per_step'), "e2e", v.get('e2e'))
    else: print(k, v)
PY
