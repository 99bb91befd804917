for sp in 1 2 4; do echo "== D2H split $sp"; CWTB_D2H_SPLIT=$sp timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e',d['e2e'])"; done
