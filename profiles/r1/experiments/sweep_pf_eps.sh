run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step %.4f launches %d'%(d['ms_per_step'],d['launches_per_step']))"; }
run A=1
run CWTB_PF_DIST=74
run CWTB_PF_DIST=296
run CWTB_PF_DIST=444
run CWTB_PF_DIST_A=74
run CWTB_PF_DIST_A=296
run CWTB_PF_DIST_A=0
run CWTB_BAND_EPS=1e-16
run CWTB_BAND_EPS=1e-13
run CWTB_DIRECT_MAX=12
run CWTB_CHAINS=3
run A=2
