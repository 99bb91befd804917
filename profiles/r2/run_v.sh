mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for p in 0 1; do echo "#### CWTB_BATCH_PIPELINE=$p"; CWTB_BATCH_PIPELINE=$p timeout 600 python bench.py --configs 5 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['5']; print('config2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'resident', d['e2e']['resident']['ms_per_step'], '| config5', c['ms_per_step'], 'e2e', c['e2e'])
"; done
} | tee gpurun_out/sweep_v.txt
