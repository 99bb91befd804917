mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
pa=sum(v[1] for n,v in k.items() if n.startswith('PassABody') and not n.endswith('-1>'))
pb=sum(v[1] for n,v in k.items() if n.startswith('PassBBody') and ', -1' not in n)
print('ms_per_step %.4f launches %d passA %.4f passB %.4f'%(d['ms_per_step'],d['launches_per_step'],pa,pb))"; }
run A=1
run CWTB_PASSB_REV=0
run CWTB_K2_512_MAX=19
run CWTB_K2_512_MAX=15
run CWTB_K2_512_MAX=17
run CWTB_PASSB_REV=0 CWTB_K2_512_MAX=19
run A=2
timeout 300 python profiles/micro/resident_breakdown.py 2>&1 | grep -v Warn | tail -8
