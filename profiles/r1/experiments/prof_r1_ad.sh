timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
print('ms_per_step %.3f  '%d['ms_per_step']+'  '.join('%s=%.3f'%(k.replace('Body<double, ','<').replace(', 1, 1>','b>').replace(', 0, 1>','d>'),v[1]) for k,v in sorted(ks.items(),key=lambda kv:-kv[1][1])[:5]))
"; }
run CWTB_GAUSS_REC=0
run CWTB_GAUSS_REC=1
