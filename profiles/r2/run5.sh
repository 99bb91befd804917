mkdir -p gpurun_out /tmp/prof
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-2300; }
run CWTB_NONE=1
run CWTB_FUSED=2
run CWTB_FUSED=2 CWTB_RING=3 CWTB_AHEAD=1
run CWTB_FUSED=2 CWTB_RING=5 CWTB_AHEAD=2
run CWTB_FUSED=1
timeout 900 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_fullsize.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12
