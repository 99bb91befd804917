cp pycwt_b200/libcwtb200.so /tmp/lib_orig.so
for v in V0 V1 V2 V3 V4 V5; do
  cp build/variants/lib_$v.so pycwt_b200/libcwtb200.so
  echo "=== variant $v"
  timeout 300 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden or plan_classes" 2>&1 | tail -1
  timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f'%d['ms_per_step'])
ks=d['kernels']
print('  '+'  '.join('%s=%.3f'%(k.replace('Body<double, ','<').replace(', 1, 1>','b>').replace(', 0, 1>','d>'),v[1]) for k,v in sorted(ks.items(),key=lambda kv:-kv[1][1])[:18]))
"
done
cp /tmp/lib_orig.so pycwt_b200/libcwtb200.so
