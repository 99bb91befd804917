timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-100; }
run CWTB_STREAMS=1
run CWTB_STREAMS=2
run CWTB_STREAMS=3
run CWTB_STREAMS=2
run CWTB_STREAMS=3
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_cwt.py -x -q -k "plan_classes or chirp32k" 2>&1 | tail -3
