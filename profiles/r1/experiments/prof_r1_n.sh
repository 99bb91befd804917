timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -1
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
print('ms_per_step %.3f  '%d['ms_per_step']+'  '.join('%s=%.3f'%(k.replace('Body<double, ','<').replace(', 1, 1>','b>').replace(', 0, 1>','d>'),v[1]) for k,v in sorted(ks.items(),key=lambda kv:-kv[1][1])[:9]))
"; }
run CWTB_PF_DIST=32
run CWTB_PF_DIST=64
run CWTB_PF_DIST=100
run CWTB_PF_DIST=148
run CWTB_PF_DIST=200
run CWTB_PF_DIST=148 CWTB_PF_DIST_A=64
run CWTB_PF_DIST=148 CWTB_PF_DIST_A=148
run CWTB_PF_DIST=148 CWTB_PF_DIST_A=300
