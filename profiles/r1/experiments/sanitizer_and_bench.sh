bash profiles/run_r1_sanitizer.sh
timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-120
