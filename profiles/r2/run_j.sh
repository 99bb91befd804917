mkdir -p gpurun_out
{
echo "#### bench.py --configs 5"; timeout 600 python bench.py --configs 5 --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['5']; print('config2', d['ms_per_step'], 'config5', c['ms_per_step'])
for k,v in c['roofline']['kernels'].items():
    if v['ms']>0.15: print('   ',k,v)
"
echo "#### script 5"; timeout 300 python profiles/micro/config_kernels.py 5 --prof 2>&1 | grep -v "coarse:\|fwd:"
echo "#### script 2,3,4,5"; timeout 300 python profiles/micro/config_kernels.py 2,3,4,5 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
} | tee gpurun_out/cmp_j.txt
