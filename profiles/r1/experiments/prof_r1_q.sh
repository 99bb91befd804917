timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -1
timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
print('ms_per_step %.3f  '%d['ms_per_step']+'  '.join('%s=%.3f'%(k.replace('Body<double, ','<').replace(', 1, 1>','b>').replace(', 0, 1>','d>'),v[1]) for k,v in sorted(ks.items(),key=lambda kv:-kv[1][1])[:8]))
"
# sanitizers on a small but representative subset (every kernel type)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden and (nino3_morlet_tutorial or chirp4000_paul or chirp32k) or plan_classes or fp32" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_xwt_wct.py -x -q -k "wct_golden or smooth or seeded_exact or batch" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "chirp4000_morlet or plan_classes" 2>&1 | tail -4
