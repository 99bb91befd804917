run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
print('ms_per_step %.3f  launches %d  '%(d['ms_per_step'],d['launches_per_step'])+'  '.join('%s=%.3f'%(k.replace('Body<double, ','<').replace(', 1, 1>','b>').replace(', 0, 1>','d>')[:30],v[1]) for k,v in sorted(ks.items(),key=lambda kv:-kv[1][1])[:5]))
"; }
run CWTB_GROUP=1
run CWTB_GROUP=2
run CWTB_GROUP=3
run CWTB_GROUP=4
run CWTB_GROUP=8
run CWTB_GROUP=16
run CWTB_GROUP=32
run CWTB_FUSED=1 CWTB_RING=3
run CWTB_FUSED=1 CWTB_RING=2
