timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-110
timeout 300 python profiles/other_configs.py 2>&1 | grep "config3\|config4\|config5 slice"
