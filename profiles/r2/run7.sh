mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xwt_wct.py tests/test_gpu_cwt.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python profiles/misc_breakdown.py 2>&1 | grep -v Warning | tee gpurun_out/misc_breakdown_r2_a.txt
timeout 600 python bench.py --steps 10 --warmup 3 --configs 2,3,5 > gpurun_out/bench_r2_try2.json 2> gpurun_out/bench_r2_try2.err; tail -c 300 gpurun_out/bench_r2_try2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_try2.json').read().strip().splitlines()[-1])
print("c2 ms %.3f e2e %.1f ms resident %.2f ms" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['resident']['ms_per_step']))
for f in ('paul','dog'):
    r=d['configs']['3'][f]; print(f, "ms %.4f" % r['ms_per_step'], "frac step %.3f dom %s %.3f" % (r['roofline']['step']['frac'], r['roofline'].get('kernel'), r['roofline'].get('frac',0)))
r=d['configs']['5']; print("c5 ms %.3f step frac %.3f dom %s %.3f e2e %.1f ms" % (r['ms_per_step'], r['roofline']['step']['frac'], r['roofline'].get('kernel'), r['roofline'].get('frac',0), r['e2e']['ms_per_step']))
PY
