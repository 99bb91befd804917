cp pycwt_b200/libcwtb200.so /tmp/lib_orig.so
for v in FA FB FC FD FE; do
  cp build/variants/lib_$v.so pycwt_b200/libcwtb200.so
  echo "=== variant $v"
  timeout 300 python -m pytest tests/test_gpu_cwt.py -x -q -k "fp32 or f32" 2>&1 | tail -1
  timeout 300 python profiles/other_configs.py 2>&1 | grep "config3\|config5 slice"
done
cp /tmp/lib_orig.so pycwt_b200/libcwtb200.so
