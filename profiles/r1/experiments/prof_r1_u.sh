timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -1
timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-120
timeout 600 python profiles/other_configs.py 2>&1 | grep -v Warning | tail -18
