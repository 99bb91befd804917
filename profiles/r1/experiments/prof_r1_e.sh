mkdir -p gpurun_out
cp pycwt_b200/libcwtb200.so /tmp/lib_orig.so
for v in A B C D E F; do
  cp build/variants/lib_$v.so pycwt_b200/libcwtb200.so
  echo "=== variant $v"
  timeout 300 python -m pytest tests/test_gpu_cwt.py -x -q -k "golden or plan_classes or fp32" 2>&1 | tail -1
  CWTB_GROUP=32 CWTB_L2_PERSIST=0 timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f'%d['ms_per_step'])
for k,v in sorted(d['kernels'].items(),key=lambda kv:-kv[1][1])[:14]:
    print('   %-34s ms %.4f rows %4d us/row %.2f'%(k,v[1],v[2],1e3*v[1]/max(v[2],1)))
"
done
cp /tmp/lib_orig.so pycwt_b200/libcwtb200.so
