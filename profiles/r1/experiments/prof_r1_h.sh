mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f launches %d'%(d['ms_per_step'],d['launches_per_step']))
for k,v in sorted(d['kernels'].items(),key=lambda kv:-kv[1][1])[:16]:
    print('   %-40s ms %.4f rows %4d us/row %.2f'%(k[:40],v[1],v[2],1e3*v[1]/max(v[2],1)))
"; }
run CWTB_FUSED=0
run CWTB_FUSED=1 CWTB_RING=3
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])"
