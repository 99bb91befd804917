mkdir -p gpurun_out
{
for p in 1; do echo "#### CWTB_BATCH_PIPELINE=$p"; CWTB_BATCH_PIPELINE=$p timeout 600 python bench.py --configs 5 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['5']; print('config2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'resident', d['e2e']['resident']['ms_per_step'], '| config5', c['ms_per_step'], 'e2e', c['e2e'])
"; done
CWTB_BATCH_MB=16384 timeout 600 python bench.py --configs 5 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['5']; print('BATCH_MB=16384 config5 e2e', c['e2e'])
"
timeout 600 python -m pytest tests/test_gpu_xwt_wct.py tests/test_gpu_fullsize.py -x -q -m gpu -k "batch or config5" 2>&1 | tail -2
} | tee gpurun_out/sweep_w.txt
