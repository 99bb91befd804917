timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step %.4f e2e %.2f ms %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['value']))"
timeout 300 python profiles/micro/e2e_breakdown.py 2>&1 | grep -v Warn | tail -9
