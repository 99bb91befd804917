mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-1800; }
run CWTB_NONE=1
run CWTB_EXPAND_MIN_R=2
echo "== full bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_try1.json 2> gpurun_out/bench_r2_try1.err; tail -c 600 gpurun_out/bench_r2_try1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_try1.json').read().strip().splitlines()[-1])
print("value %.3e ms %.3f e2e %.3e (%.1f ms) resident %.1f ms" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['resident']['ms_per_step']))
r=d['roofline']; print("dominant", r['kernel'], "frac %.3f"%r['frac'], "step frac %.3f"%r['step']['frac'])
print("cpu", d['cpu_baseline'])
for k,v in d['configs'].items():
    print(k, json.dumps(v)[:1500])
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-700
