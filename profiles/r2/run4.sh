mkdir -p gpurun_out /tmp/prof
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-2300; }
cp pycwt_b200/libcwtb200.so /tmp/lib_main.so
run CWTB_NONE=1
run CWTB_FUSED=2
run CWTB_FUSED=2 CWTB_RING=3 CWTB_AHEAD=1
run CWTB_FUSED=2 CWTB_RING=5 CWTB_AHEAD=2
run CWTB_FUSED=2 CWTB_RING=5 CWTB_AHEAD=3
run CWTB_FUSED=1
timeout 300 python -m pytest tests/test_gpu_cwt.py -x -q -m gpu 2>&1 | tail -2
CWTB_FUSED=2 timeout 300 python -m pytest tests/test_gpu_cwt.py -x -q -m gpu 2>&1 | tail -2
cp build/variants/lib_xch2.so pycwt_b200/libcwtb200.so; echo "variant: 2 chains"; run CWTB_NONE=1
cp build/variants/lib_xseg2.so pycwt_b200/libcwtb200.so; echo "variant: 2 segments"; run CWTB_NONE=1
cp /tmp/lib_main.so pycwt_b200/libcwtb200.so
