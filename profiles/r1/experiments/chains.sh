run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step %.4f launches %d'%(d['ms_per_step'],d['launches_per_step']))"; }
timeout 600 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden or plan_classes or linearity" 2>&1 | tail -1
run CWTB_CHAINS=2
run CWTB_CHAINS=3
run CWTB_CHAINS=4
run CWTB_CHAINS=1
run CWTB_CHAINS=3 CWTB_GROUP=8
run CWTB_CHAINS=4 CWTB_GROUP=8
run CWTB_CHAINS=2
