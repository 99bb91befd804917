# does keeping Z in L2 pay now?  small launch groups / fused kernel / persisting window
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
pa=sum(v[1] for n,v in k.items() if n.startswith('PassABody') and not n.endswith('-1>'))
pb=sum(v[1] for n,v in k.items() if n.startswith('PassBBody') and ', -1' not in n)
print('ms_per_step %.4f launches %d passA %.4f passB %.4f'%(d['ms_per_step'],d['launches_per_step'],pa,pb))"; }
run A=1
run CWTB_GROUP=2
run CWTB_GROUP=4
run CWTB_GROUP=8
run CWTB_GROUP=4 CWTB_L2_PERSIST=1
run CWTB_FUSED=1
run CWTB_FUSED=1 CWTB_RING=4
run CWTB_K2_BAND=9
run CWTB_K2_BAND=9 CWTB_GROUP=4
